// h2.h -- the sliver of HTTP/2 (RFC 9113) and HPACK (RFC 7541) a unary gRPC call needs, for the
// native load generator's gRPC transport (C4 / C5 of BASELINE.json: ModelInfer with
// raw_input_contents, src/python/library/tritonclient/grpc/_client.py:1445-1572 rides on
// grpcio's C core for this) and for the canned-response gRPC stub server.
//
// Client side: connection preface + SETTINGS, one stream per request (HEADERS with a constant
// header block, DATA frames carrying the 5-byte gRPC message prefix + the protobuf bytes),
// SETTINGS / PING acknowledgements, flow-control bookkeeping.  Header blocks we SEND use only
// literal-without-indexing fields and two static-table indices, so no encoder state exists;
// header blocks we RECEIVE are skipped (a response counts as successful when a DATA frame
// arrived before the stream ended -- gRPC reports errors as trailers-only responses), so no
// decoder state is needed either.  (The C++ gRPC front end and the servers do decode: HpackDecoder
// below, dynamic table and Huffman strings included.)
#ifndef TB200_CSRC_H2_H_
#define TB200_CSRC_H2_H_

#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <utility>
#include <vector>

#include "hpack_huffman.h"

namespace tb200 { namespace h2 {

enum FrameType : uint8_t { DATA = 0, HEADERS = 1, RST_STREAM = 3, SETTINGS = 4, PING = 6, GOAWAY = 7, WINDOW_UPDATE = 8, CONTINUATION = 9 };
constexpr uint8_t kEndStream = 0x1, kAck = 0x1, kEndHeaders = 0x4, kPadded = 0x8, kPriority = 0x20;
constexpr uint32_t kDefaultWindow = 65535, kDefaultMaxFrame = 16384;
constexpr uint16_t kSettingsEnablePush = 2, kSettingsInitialWindow = 4, kSettingsMaxFrame = 5;

inline void put_frame_header(uint8_t* out, uint32_t length, uint8_t type, uint8_t flags, uint32_t stream) {
  out[0] = static_cast<uint8_t>(length >> 16);
  out[1] = static_cast<uint8_t>(length >> 8);
  out[2] = static_cast<uint8_t>(length);
  out[3] = type;
  out[4] = flags;
  out[5] = static_cast<uint8_t>((stream >> 24) & 0x7F);
  out[6] = static_cast<uint8_t>(stream >> 16);
  out[7] = static_cast<uint8_t>(stream >> 8);
  out[8] = static_cast<uint8_t>(stream);
}
inline std::string frame(uint8_t type, uint8_t flags, uint32_t stream, const std::string& payload) {
  std::string f(9, '\0');
  put_frame_header(reinterpret_cast<uint8_t*>(&f[0]), static_cast<uint32_t>(payload.size()), type, flags, stream);
  return f + payload;
}
inline void put_u32(std::string* s, uint32_t v) {
  const char b[4] = {static_cast<char>(v >> 24), static_cast<char>(v >> 16), static_cast<char>(v >> 8), static_cast<char>(v)};
  s->append(b, 4);
}
inline std::string setting(uint16_t id, uint32_t value) {
  std::string s;
  s.push_back(static_cast<char>(id >> 8));
  s.push_back(static_cast<char>(id));
  put_u32(&s, value);
  return s;
}
inline std::string window_update(uint32_t stream, uint32_t increment) {
  std::string p;
  put_u32(&p, increment & 0x7FFFFFFFu);
  return frame(WINDOW_UPDATE, 0, stream, p);
}

// HPACK integer with an N-bit prefix (RFC 7541 5.1); `first` carries the pattern bits above it
inline void hpack_int(std::string* out, uint8_t first, int prefix_bits, size_t value) {
  const size_t limit = (1u << prefix_bits) - 1;
  if (value < limit) {
    out->push_back(static_cast<char>(first | value));
    return;
  }
  out->push_back(static_cast<char>(first | limit));
  value -= limit;
  while (value >= 128) {
    out->push_back(static_cast<char>((value & 0x7F) | 0x80));
    value >>= 7;
  }
  out->push_back(static_cast<char>(value));
}
// literal header field without indexing, new name, raw (non-Huffman) strings (RFC 7541 6.2.2)
inline void hpack_literal(std::string* out, const std::string& name, const std::string& value) {
  out->push_back('\0');
  hpack_int(out, 0x00, 7, name.size());
  out->append(name);
  hpack_int(out, 0x00, 7, value.size());
  out->append(value);
}

// what a client sends first: preface, SETTINGS (no push, 1 MiB stream windows), and a
// connection window large enough that responses never wait for us
constexpr uint32_t kOurStreamWindow = 1u << 20;
constexpr uint32_t kOurConnWindow = 1u << 30;
inline std::string client_preface() {
  std::string s = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";
  s += frame(SETTINGS, 0, 0, setting(kSettingsEnablePush, 0) + setting(kSettingsInitialWindow, kOurStreamWindow));
  s += window_update(0, kOurConnWindow - kDefaultWindow);
  return s;
}
// request header block of a unary gRPC call
inline std::string grpc_request_headers(const std::string& authority, const std::string& path) {
  std::string b;
  b.push_back(static_cast<char>(0x83));  // :method: POST   (static table index 3)
  b.push_back(static_cast<char>(0x86));  // :scheme: http   (static table index 6)
  hpack_literal(&b, ":path", path);
  hpack_literal(&b, ":authority", authority);
  hpack_literal(&b, "content-type", "application/grpc");
  hpack_literal(&b, "te", "trailers");
  hpack_literal(&b, "user-agent", "tb200-loadgen/1.0");
  return b;
}
// response header block and trailers of a successful call (server side)
inline std::string grpc_response_headers() {
  std::string b;
  b.push_back(static_cast<char>(0x88));  // :status: 200    (static table index 8)
  hpack_literal(&b, "content-type", "application/grpc");
  return b;
}
inline std::string grpc_trailers_ok() {
  std::string b;
  hpack_literal(&b, "grpc-status", "0");
  return b;
}
// 5-byte gRPC length-prefixed-message header (uncompressed)
inline void put_grpc_prefix(uint8_t* out, uint32_t message_bytes) {
  out[0] = 0;
  out[1] = static_cast<uint8_t>(message_bytes >> 24);
  out[2] = static_cast<uint8_t>(message_bytes >> 16);
  out[3] = static_cast<uint8_t>(message_bytes >> 8);
  out[4] = static_cast<uint8_t>(message_bytes);
}

struct FrameView {
  uint32_t length = 0;
  uint8_t type = 0, flags = 0;
  uint32_t stream = 0;
  const uint8_t* payload = nullptr;
};
// the frame at buf[0..n) if it is complete; returns its total size (9 + length) or 0
inline size_t parse_frame(const uint8_t* buf, size_t n, FrameView* f) {
  if (n < 9) return 0;
  const uint32_t len = (static_cast<uint32_t>(buf[0]) << 16) | (static_cast<uint32_t>(buf[1]) << 8) | buf[2];
  if (n < 9 + static_cast<size_t>(len)) return 0;
  f->length = len;
  f->type = buf[3];
  f->flags = buf[4];
  f->stream = ((static_cast<uint32_t>(buf[5]) & 0x7F) << 24) | (static_cast<uint32_t>(buf[6]) << 16) | (static_cast<uint32_t>(buf[7]) << 8) | buf[8];
  f->payload = buf + 9;
  return 9 + static_cast<size_t>(len);
}
inline uint32_t get_u32(const uint8_t* p) {
  return (static_cast<uint32_t>(p[0]) << 24) | (static_cast<uint32_t>(p[1]) << 16) | (static_cast<uint32_t>(p[2]) << 8) | p[3];
}

// ---- HPACK Huffman decoding (RFC 7541 5.2, code table in hpack_huffman.h) --------------------
// A binary trie over the 257 codes, built once; decoding walks it bit by bit.  The string must
// end on a symbol boundary followed by fewer than 8 padding bits that are all ones, and must not
// contain EOS.
class HuffmanDecoder {
 public:
  static bool Decode(const uint8_t* p, size_t n, std::string* out) {
    static const HuffmanDecoder kTrie;
    out->clear();
    int node = 0;
    int since_symbol = 0;   // bits consumed since the last complete symbol
    bool all_ones = true;   // ... and whether they were all ones
    for (size_t i = 0; i < n; ++i) {
      for (int bit = 7; bit >= 0; --bit) {
        const int b = (p[i] >> bit) & 1;
        node = kTrie.nodes_[static_cast<size_t>(node)].child[b];
        if (node < 0) return false;
        ++since_symbol;
        all_ones = all_ones && b == 1;
        const int sym = kTrie.nodes_[static_cast<size_t>(node)].symbol;
        if (sym >= 0) {
          if (sym == 256) return false;  // EOS inside a string is a decoding error
          out->push_back(static_cast<char>(sym));
          node = 0;
          since_symbol = 0;
          all_ones = true;
        }
      }
    }
    return since_symbol < 8 && all_ones;
  }

 private:
  struct Node {
    int child[2] = {-1, -1};
    int symbol = -1;
  };
  HuffmanDecoder() {
    nodes_.emplace_back();
    for (int s = 0; s < 257; ++s) {
      int node = 0;
      for (int bit = kHpackHuffman[s].bits - 1; bit >= 0; --bit) {
        const int b = (kHpackHuffman[s].code >> bit) & 1;
        if (nodes_[static_cast<size_t>(node)].child[b] < 0) {
          nodes_[static_cast<size_t>(node)].child[b] = static_cast<int>(nodes_.size());
          nodes_.emplace_back();
        }
        node = nodes_[static_cast<size_t>(node)].child[b];
      }
      nodes_[static_cast<size_t>(node)].symbol = s;
    }
  }
  std::vector<Node> nodes_;
};

// ---- HPACK decoding (the C++ gRPC front end reads grpc-status / grpc-message, the servers :path)
// Static table + dynamic table + integer / string literals, raw or Huffman coded
// (RFC 7541 2.3, 5, 6).
class HpackDecoder {
 public:
  using Field = std::pair<std::string, std::string>;

  bool Decode(const uint8_t* p, size_t n, std::vector<Field>* out) {
    const uint8_t* end = p + n;
    while (p < end) {
      const uint8_t b = *p;
      if (b & 0x80) {  // indexed field
        uint64_t index;
        if (!Int(&p, end, 7, &index) || index == 0) return false;
        Field f;
        if (!Lookup(index, &f)) return false;
        out->push_back(std::move(f));
      } else if ((b & 0xE0) == 0x20) {  // dynamic table size update
        uint64_t size;
        if (!Int(&p, end, 5, &size)) return false;
        max_size_ = size;
        Evict();
      } else {
        const bool add = (b & 0xC0) == 0x40;  // literal with incremental indexing
        uint64_t index;
        if (!Int(&p, end, add ? 6 : 4, &index)) return false;
        Field f;
        if (index != 0) {
          Field named;
          if (!Lookup(index, &named)) return false;
          f.first = named.first;
        } else if (!Str(&p, end, &f.first)) {
          return false;
        }
        if (!Str(&p, end, &f.second)) return false;
        if (add) {
          size_ += f.first.size() + f.second.size() + 32;
          table_.push_front(f);
          Evict();
        }
        out->push_back(std::move(f));
      }
    }
    return true;
  }

 private:
  static bool Int(const uint8_t** p, const uint8_t* end, int prefix_bits, uint64_t* v) {
    if (*p >= end) return false;
    const uint64_t limit = (1u << prefix_bits) - 1;
    *v = *(*p)++ & limit;
    if (*v < limit) return true;
    for (int shift = 0; shift < 56; shift += 7) {
      if (*p >= end) return false;
      const uint8_t b = *(*p)++;
      *v += static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return true;
    }
    return false;
  }
  bool Str(const uint8_t** p, const uint8_t* end, std::string* s) {
    if (*p >= end) return false;
    const bool huffman = (**p & 0x80) != 0;
    uint64_t len;
    if (!Int(p, end, 7, &len) || len > static_cast<uint64_t>(end - *p)) return false;
    if (huffman) {
      if (!HuffmanDecoder::Decode(*p, static_cast<size_t>(len), s)) return false;
    } else {
      s->assign(reinterpret_cast<const char*>(*p), len);
    }
    *p += len;
    return true;
  }
  bool Lookup(uint64_t index, Field* f) const {
    static const char* const kStatic[61][2] = {
        {":authority", ""}, {":method", "GET"}, {":method", "POST"}, {":path", "/"}, {":path", "/index.html"},
        {":scheme", "http"}, {":scheme", "https"}, {":status", "200"}, {":status", "204"}, {":status", "206"},
        {":status", "304"}, {":status", "400"}, {":status", "404"}, {":status", "500"}, {"accept-charset", ""},
        {"accept-encoding", "gzip, deflate"}, {"accept-language", ""}, {"accept-ranges", ""}, {"accept", ""},
        {"access-control-allow-origin", ""}, {"age", ""}, {"allow", ""}, {"authorization", ""}, {"cache-control", ""},
        {"content-disposition", ""}, {"content-encoding", ""}, {"content-language", ""}, {"content-length", ""},
        {"content-location", ""}, {"content-range", ""}, {"content-type", ""}, {"cookie", ""}, {"date", ""}, {"etag", ""},
        {"expect", ""}, {"expires", ""}, {"from", ""}, {"host", ""}, {"if-match", ""}, {"if-modified-since", ""},
        {"if-none-match", ""}, {"if-range", ""}, {"if-unmodified-since", ""}, {"last-modified", ""}, {"link", ""},
        {"location", ""}, {"max-forwards", ""}, {"proxy-authenticate", ""}, {"proxy-authorization", ""}, {"range", ""},
        {"referer", ""}, {"refresh", ""}, {"retry-after", ""}, {"server", ""}, {"set-cookie", ""},
        {"strict-transport-security", ""}, {"transfer-encoding", ""}, {"user-agent", ""}, {"vary", ""}, {"via", ""},
        {"www-authenticate", ""}};
    if (index <= 61) {
      f->first = kStatic[index - 1][0];
      f->second = kStatic[index - 1][1];
      return true;
    }
    if (index - 62 >= table_.size()) return false;
    *f = table_[index - 62];
    return true;
  }
  void Evict() {
    while (size_ > max_size_ && !table_.empty()) {
      size_ -= table_.back().first.size() + table_.back().second.size() + 32;
      table_.pop_back();
    }
  }
  std::deque<Field> table_;
  size_t size_ = 0, max_size_ = 4096;
};

}}  // namespace tb200::h2

#endif  // TB200_CSRC_H2_H_
