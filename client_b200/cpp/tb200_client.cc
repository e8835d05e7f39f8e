// tb200_client.cc -- implementation of the C++ front end (see tb200_client.h).
//
// Behaviour restated from the reference C++ client: scatter-list inputs
// (src/c++/library/common.cc:112-289), request JSON layout (http_client.cc:412-578),
// binary<->JSON tensor conversions (:581-678, :1156-1281), response splitting (:1043-1136),
// statistics (common.cc:56-106).  The transport is our own socket code.
#include "tb200_client.h"

#include <zlib.h>

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <iostream>
#include <limits>

#include "../../include/tb200.h"
#include "json.h"

namespace tb200 { namespace client {

using json::Value;

const Error Error::Success("");

std::ostream& operator<<(std::ostream& out, const Error& err) {
  if (!err.msg_.empty()) out << err.msg_;
  return out;
}

// ---- RequestTimers / statistics ---------------------------------------------------------

uint64_t RequestTimers::CaptureTimestamp(Kind kind) {
  uint64_t& ts = timestamps_[static_cast<size_t>(kind)];
  ts = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(
                                 std::chrono::high_resolution_clock::now().time_since_epoch())
                                 .count());
  return ts;
}

uint64_t RequestTimers::Duration(Kind start, Kind end) const {
  const uint64_t s = timestamps_[static_cast<size_t>(start)], e = timestamps_[static_cast<size_t>(end)];
  if (s == 0 || e == 0 || s > e) return std::numeric_limits<uint64_t>::max();
  return e - s;
}

Error InferenceServerClient::ClientInferStat(InferStat* infer_stat) const {
  std::lock_guard<std::mutex> lk(stat_mu_);
  *infer_stat = infer_stat_;
  return Error::Success;
}

Error InferenceServerClient::UpdateInferStat(const RequestTimers& timer) {
  using K = RequestTimers::Kind;
  const uint64_t bad = std::numeric_limits<uint64_t>::max();
  const uint64_t total = timer.Duration(K::REQUEST_START, K::REQUEST_END);
  const uint64_t send = timer.Duration(K::SEND_START, K::SEND_END);
  const uint64_t recv = timer.Duration(K::RECV_START, K::RECV_END);
  if (total == bad || send == bad || recv == bad) {
    auto span = [&](const char* what, K a, K b) {
      return timer.Timestamp(a) > timer.Timestamp(b)
                 ? std::string(" ") + what + " time from " + std::to_string(timer.Timestamp(a)) + " to " +
                       std::to_string(timer.Timestamp(b)) + "."
                 : std::string();
    };
    return Error("Timer not set correctly." + span("Request", K::REQUEST_START, K::REQUEST_END) +
                 span("Send", K::SEND_START, K::SEND_END) + span("Receive", K::RECV_START, K::RECV_END));
  }
  std::lock_guard<std::mutex> lk(stat_mu_);
  infer_stat_.completed_request_count++;
  infer_stat_.cumulative_total_request_time_ns += total;
  infer_stat_.cumulative_send_time_ns += send;
  infer_stat_.cumulative_receive_time_ns += recv;
  return Error::Success;
}

// ---- InferInput ------------------------------------------------------------------------

InferInput::InferInput(const std::string& name, const std::vector<int64_t>& dims, const std::string& datatype)
    : name_(name), shape_(dims), datatype_(datatype) {}

Error InferInput::Create(InferInput** infer_input, const std::string& name, const std::vector<int64_t>& dims,
                         const std::string& datatype) {
  *infer_input = new InferInput(name, dims, datatype);
  return Error::Success;
}

Error InferInput::SetShape(const std::vector<int64_t>& dims) {
  shape_ = dims;
  return Error::Success;
}

Error InferInput::Reset() {
  bufs_.clear();
  buf_byte_sizes_.clear();
  str_bufs_.clear();
  bufs_idx_ = 0;
  byte_size_ = 0;
  io_type_ = NONE;
  return Error::Success;
}

Error InferInput::AppendRaw(const std::vector<uint8_t>& input) { return AppendRaw(input.data(), input.size()); }

Error InferInput::AppendRaw(const uint8_t* input, size_t input_byte_size) {
  // borrowed: the caller keeps the buffer alive until the request completed (common.h:274-293)
  byte_size_ += input_byte_size;
  bufs_.push_back(input);
  buf_byte_sizes_.push_back(input_byte_size);
  io_type_ = RAW;
  return Error::Success;
}

Error InferInput::SetSharedMemory(const std::string& name, size_t byte_size, size_t offset) {
  shm_name_ = name;
  shm_offset_ = offset;
  byte_size_ = byte_size;
  io_type_ = SHARED_MEMORY;
  return Error::Success;
}

Error InferInput::SharedMemoryInfo(std::string* name, size_t* byte_size, size_t* offset) const {
  if (io_type_ != SHARED_MEMORY) return Error("The input has not been set with the shared memory.");
  *name = shm_name_;
  *offset = shm_offset_;
  *byte_size = byte_size_;
  return Error::Success;
}

Error InferInput::AppendFromString(const std::vector<std::string>& input) {
  // BYTES framing: <u32 little-endian length><payload>, no terminator (common.cc:169-183)
  str_bufs_.emplace_back();
  std::string& framed = str_bufs_.back();
  size_t total = 0;
  for (const std::string& s : input) total += sizeof(uint32_t) + s.size();
  framed.reserve(total);
  for (const std::string& s : input) {
    const uint32_t len = static_cast<uint32_t>(s.size());
    framed.append(reinterpret_cast<const char*>(&len), sizeof(len));
    framed.append(s);
  }
  return AppendRaw(reinterpret_cast<const uint8_t*>(framed.data()), framed.size());
}

Error InferInput::RawData(const uint8_t** buf, size_t* byte_size) {
  if (!bufs_.empty()) {
    *buf = bufs_[0];
    *byte_size = buf_byte_sizes_[0];
  } else {
    *buf = nullptr;
    *byte_size = 0;
  }
  return Error::Success;
}

Error InferInput::ByteSize(size_t* byte_size) const {
  *byte_size = byte_size_;
  return Error::Success;
}

Error InferInput::SetBinaryData(const bool binary_data) {
  binary_data_ = binary_data;
  return Error::Success;
}

Error InferInput::PrepareForRequest() {
  bufs_idx_ = 0;
  buf_pos_ = 0;
  return Error::Success;
}

Error InferInput::GetNext(uint8_t* buf, size_t size, size_t* input_bytes, bool* end_of_input) {
  size_t copied = 0;
  while (bufs_idx_ < bufs_.size() && size > 0) {
    const size_t left = buf_byte_sizes_[bufs_idx_] - buf_pos_;
    const size_t n = std::min(left, size);
    if (n > 0) {
      memcpy(buf, bufs_[bufs_idx_] + buf_pos_, n);
      buf += n;
      size -= n;
      buf_pos_ += n;
      copied += n;
    }
    if (buf_pos_ == buf_byte_sizes_[bufs_idx_]) {
      ++bufs_idx_;
      buf_pos_ = 0;
    }
  }
  *input_bytes = copied;
  *end_of_input = bufs_idx_ >= bufs_.size();
  return Error::Success;
}

Error InferInput::GetNext(const uint8_t** buf, size_t* input_bytes, bool* end_of_input) {
  if (bufs_idx_ < bufs_.size()) {
    *buf = bufs_[bufs_idx_];
    *input_bytes = buf_byte_sizes_[bufs_idx_];
    ++bufs_idx_;
  } else {
    *buf = nullptr;
    *input_bytes = 0;
  }
  *end_of_input = bufs_idx_ >= bufs_.size();
  return Error::Success;
}

// ---- InferRequestedOutput --------------------------------------------------------------

InferRequestedOutput::InferRequestedOutput(const std::string& name, const std::string& datatype, const size_t class_count)
    : name_(name), datatype_(datatype), class_count_(class_count) {}

Error InferRequestedOutput::Create(InferRequestedOutput** infer_output, const std::string& name, const size_t class_count,
                                   const std::string& datatype) {
  *infer_output = new InferRequestedOutput(name, datatype, class_count);
  return Error::Success;
}

Error InferRequestedOutput::SetSharedMemory(const std::string& region_name, const size_t byte_size, const size_t offset) {
  shm_name_ = region_name;
  shm_byte_size_ = byte_size;
  shm_offset_ = offset;
  io_type_ = SHARED_MEMORY;
  return Error::Success;
}

Error InferRequestedOutput::UnsetSharedMemory() {
  shm_name_.clear();
  shm_byte_size_ = 0;
  shm_offset_ = 0;
  io_type_ = NONE;
  return Error::Success;
}

Error InferRequestedOutput::SharedMemoryInfo(std::string* name, size_t* byte_size, size_t* offset) const {
  if (io_type_ != SHARED_MEMORY) return Error("The input has not been set with the shared memory.");
  *name = shm_name_;
  *offset = shm_offset_;
  *byte_size = shm_byte_size_;
  return Error::Success;
}

Error InferRequestedOutput::SetBinaryData(const bool binary_data) {
  binary_data_ = binary_data;
  return Error::Success;
}

// ---- tensor bytes <-> JSON "data" ----------------------------------------------------------

namespace detail {

namespace {
template <typename T, typename W>
void AppendNumbers(const uint8_t* buf, size_t n, std::vector<std::string>* items) {
  for (size_t i = 0; i < n; ++i) {
    T v;
    memcpy(&v, buf + i * sizeof(T), sizeof(T));
    items->push_back(std::to_string(static_cast<W>(v)));
  }
}
}  // namespace

// One buffer of `element_count` elements -> JSON scalars as text (http_client.cc:607-678).
Error BinaryInputToJsonText(const uint8_t* buf, size_t element_count, const std::string& datatype,
                            std::vector<std::string>* items) {
  if (datatype == "BOOL") {
    for (size_t i = 0; i < element_count; ++i) items->push_back(buf[i] ? "true" : "false");
  } else if (datatype == "UINT8") {
    AppendNumbers<uint8_t, uint64_t>(buf, element_count, items);
  } else if (datatype == "UINT16") {
    AppendNumbers<uint16_t, uint64_t>(buf, element_count, items);
  } else if (datatype == "UINT32") {
    AppendNumbers<uint32_t, uint64_t>(buf, element_count, items);
  } else if (datatype == "UINT64") {
    AppendNumbers<uint64_t, uint64_t>(buf, element_count, items);
  } else if (datatype == "INT8") {
    AppendNumbers<int8_t, int64_t>(buf, element_count, items);
  } else if (datatype == "INT16") {
    AppendNumbers<int16_t, int64_t>(buf, element_count, items);
  } else if (datatype == "INT32") {
    AppendNumbers<int32_t, int64_t>(buf, element_count, items);
  } else if (datatype == "INT64") {
    AppendNumbers<int64_t, int64_t>(buf, element_count, items);
  } else if (datatype == "FP32") {
    for (size_t i = 0; i < element_count; ++i) {
      float v;
      memcpy(&v, buf + i * sizeof(float), sizeof(float));
      std::string s;
      Value::WriteDouble(static_cast<double>(v), &s);
      items->push_back(std::move(s));
    }
  } else if (datatype == "FP64") {
    for (size_t i = 0; i < element_count; ++i) {
      double v;
      memcpy(&v, buf + i * sizeof(double), sizeof(double));
      std::string s;
      Value::WriteDouble(v, &s);
      items->push_back(std::move(s));
    }
  } else if (datatype == "BYTES") {
    size_t offset = 0;
    for (size_t i = 0; i < element_count; ++i) {
      uint32_t len;
      memcpy(&len, buf + offset, sizeof(len));
      std::string s;
      Value::WriteString(std::string(reinterpret_cast<const char*>(buf + offset + sizeof(len)), len), &s);
      items->push_back(std::move(s));
      offset += sizeof(len) + len;
    }
  } else if (datatype == "FP16" || datatype == "BF16") {
    return Error("datatype '" + datatype + "' is not supported with JSON. Please use the binary data format");
  } else {
    return Error("datatype '" + datatype + "' is invalid");
  }
  return Error::Success;
}

// Every AppendRaw buffer of the input in order, each holding prod(shape[1:]) elements
// (http_client.cc:581-604): two appended [4]-buffers of a [1,2,2] tensor flatten to 8 scalars.
Error BinaryInputsToJsonText(InferInput& input, std::vector<std::string>* items) {
  input.PrepareForRequest();
  size_t element_count = 1;
  for (size_t i = 1; i < input.Shape().size(); ++i) element_count *= static_cast<size_t>(input.Shape()[i]);
  bool end_of_input = false;
  while (!end_of_input) {
    const uint8_t* buf = nullptr;
    size_t n = 0;
    input.GetNext(&buf, &n, &end_of_input);
    if (buf != nullptr) {
      Error err = BinaryInputToJsonText(buf, element_count, input.Datatype(), items);
      if (!err.IsOk()) return err;
    }
  }
  return Error::Success;
}

namespace {

template <typename T>
void StoreInts(const Value& data, std::string* out, bool is_signed) {
  out->resize(data.size() * sizeof(T));
  for (size_t i = 0; i < data.size(); ++i) {
    T v;
    if (is_signed) {
      int64_t x = 0;
      data[i].AsInt(&x);
      v = static_cast<T>(x);
    } else {
      uint64_t x = 0;
      data[i].AsUInt(&x);
      v = static_cast<T>(x);
    }
    memcpy(&(*out)[i * sizeof(T)], &v, sizeof(T));
  }
}

// JSON "data" of a response output -> tensor bytes (http_client.cc:1156-1281).  Nested
// arrays are flattened row-major first.
void Flatten(const Value& v, Value* flat) {
  if (v.is_array()) {
    for (size_t i = 0; i < v.size(); ++i) Flatten(v[i], flat);
  } else {
    flat->Append(v);
  }
}

Error JsonOutputToBinary(const Value& data_json, const std::string& datatype, std::string* out) {
  Value data = Value::Array();
  Flatten(data_json, &data);
  const size_t n = data.size();
  if (datatype == "BOOL") {
    out->resize(n);
    for (size_t i = 0; i < n; ++i) {
      bool b = false;
      data[i].AsBool(&b);
      (*out)[i] = b ? 1 : 0;
    }
  } else if (datatype == "UINT8") {
    StoreInts<uint8_t>(data, out, false);
  } else if (datatype == "UINT16") {
    StoreInts<uint16_t>(data, out, false);
  } else if (datatype == "UINT32") {
    StoreInts<uint32_t>(data, out, false);
  } else if (datatype == "UINT64") {
    StoreInts<uint64_t>(data, out, false);
  } else if (datatype == "INT8") {
    StoreInts<int8_t>(data, out, true);
  } else if (datatype == "INT16") {
    StoreInts<int16_t>(data, out, true);
  } else if (datatype == "INT32") {
    StoreInts<int32_t>(data, out, true);
  } else if (datatype == "INT64") {
    StoreInts<int64_t>(data, out, true);
  } else if (datatype == "FP32") {
    out->resize(n * sizeof(float));
    for (size_t i = 0; i < n; ++i) {
      double d = 0.0;
      data[i].AsDouble(&d);
      const float f = static_cast<float>(d);
      memcpy(&(*out)[i * sizeof(float)], &f, sizeof(float));
    }
  } else if (datatype == "FP64") {
    out->resize(n * sizeof(double));
    for (size_t i = 0; i < n; ++i) {
      double d = 0.0;
      data[i].AsDouble(&d);
      memcpy(&(*out)[i * sizeof(double)], &d, sizeof(double));
    }
  } else if (datatype == "BYTES") {
    out->clear();
    for (size_t i = 0; i < n; ++i) {
      const std::string& s = data[i].str();
      const uint32_t len = static_cast<uint32_t>(s.size());
      out->append(reinterpret_cast<const char*>(&len), sizeof(len));
      out->append(s);
    }
  } else if (datatype == "FP16" || datatype == "BF16") {
    return Error("datatype '" + datatype + "' is not supported with JSON.");
  } else {
    return Error("datatype '" + datatype + "' is invalid");
  }
  return Error::Success;
}

std::string JoinItems(const std::vector<std::string>& items) {
  size_t total = 2;
  for (const std::string& s : items) total += s.size() + 1;
  std::string out;
  out.reserve(total);
  out.push_back('[');
  for (size_t i = 0; i < items.size(); ++i) {
    if (i) out.push_back(',');
    out.append(items[i]);
  }
  out.push_back(']');
  return out;
}

void AddShm(std::string* js, const std::string& region, size_t byte_size, size_t offset) {
  js->append("\"shared_memory_region\":");
  Value::WriteString(region, js);
  js->append(",\"shared_memory_byte_size\":" + std::to_string(byte_size));
  if (offset != 0) js->append(",\"shared_memory_offset\":" + std::to_string(offset));
}

// The inference header in the reference C++ client's member order: id, parameters,
// inputs{name,datatype,shape,parameters|data}, outputs{name,parameters} (http_client.cc:412-578).
Error BuildInferHeader(const InferOptions& options, const std::vector<InferInput*>& inputs,
                       const std::vector<const InferRequestedOutput*>& outputs, std::string* js) {
  js->clear();
  js->append("{\"id\":");
  Value::WriteString(options.request_id_, js);
  const bool has_sequence = options.sequence_id_ != 0 || !options.sequence_id_str_.empty();
  if (has_sequence || options.priority_ != 0 || options.server_timeout_ != 0 || outputs.empty()) {
    js->append(",\"parameters\":{");
    bool first = true;
    auto key = [&](const char* k) {
      if (!first) js->push_back(',');
      first = false;
      Value::WriteString(k, js);
      js->push_back(':');
    };
    if (has_sequence) {
      key("sequence_id");
      if (options.sequence_id_ != 0) js->append(std::to_string(options.sequence_id_));
      else Value::WriteString(options.sequence_id_str_, js);
      key("sequence_start");
      js->append(options.sequence_start_ ? "true" : "false");
      key("sequence_end");
      js->append(options.sequence_end_ ? "true" : "false");
    }
    if (options.priority_ != 0) {
      key("priority");
      js->append(std::to_string(options.priority_));
    }
    if (options.server_timeout_ != 0) {
      key("timeout");
      js->append(std::to_string(options.server_timeout_));
    }
    if (outputs.empty()) {
      key("binary_data_output");
      js->append("true");
    }
    for (const auto& kv : options.request_parameters) {
      const RequestParameter& p = kv.second;
      if (p.type == "string") {
        key(kv.first.c_str());
        Value::WriteString(p.value, js);
      } else if (p.type == "int") {
        key(kv.first.c_str());
        js->append(std::to_string(std::stoi(p.value)));
      } else if (p.type == "bool") {
        key(kv.first.c_str());
        js->append(p.value == "true" ? "true" : "false");
      }
    }
    js->push_back('}');
  }
  if (!inputs.empty()) {
    js->append(",\"inputs\":[");
    for (size_t i = 0; i < inputs.size(); ++i) {
      InferInput* io = inputs[i];
      if (i) js->push_back(',');
      js->append("{\"name\":");
      Value::WriteString(io->Name(), js);
      js->append(",\"datatype\":");
      Value::WriteString(io->Datatype(), js);
      js->append(",\"shape\":[");
      for (size_t d = 0; d < io->Shape().size(); ++d) {
        if (d) js->push_back(',');
        js->append(std::to_string(static_cast<uint64_t>(io->Shape()[d])));
      }
      js->push_back(']');
      if (io->IsSharedMemory()) {
        std::string region;
        size_t byte_size = 0, offset = 0;
        Error err = io->SharedMemoryInfo(&region, &byte_size, &offset);
        if (!err.IsOk()) return err;
        js->append(",\"parameters\":{");
        AddShm(js, region, byte_size, offset);
        js->push_back('}');
      } else if (io->BinaryData()) {
        size_t byte_size = 0;
        io->ByteSize(&byte_size);
        js->append(",\"parameters\":{\"binary_data_size\":" + std::to_string(byte_size) + "}");
      } else {
        std::vector<std::string> items;
        Error err = BinaryInputsToJsonText(*io, &items);
        if (!err.IsOk()) return err;
        js->append(",\"data\":" + JoinItems(items));
      }
      js->push_back('}');
    }
    js->push_back(']');
  }
  if (!outputs.empty()) {
    js->append(",\"outputs\":[");
    for (size_t i = 0; i < outputs.size(); ++i) {
      const InferRequestedOutput* io = outputs[i];
      if (i) js->push_back(',');
      js->append("{\"name\":");
      Value::WriteString(io->Name(), js);
      js->append(",\"parameters\":{");
      bool first = true;
      if (io->ClassificationCount() > 0) {
        js->append("\"classification\":" + std::to_string(io->ClassificationCount()));
        first = false;
      }
      if (io->IsSharedMemory()) {
        std::string region;
        size_t byte_size = 0, offset = 0;
        Error err = io->SharedMemoryInfo(&region, &byte_size, &offset);
        if (!err.IsOk()) return err;
        if (!first) js->push_back(',');
        AddShm(js, region, byte_size, offset);
      } else {
        if (!first) js->push_back(',');
        js->append(std::string("\"binary_data\":") + (io->BinaryData() ? "true" : "false"));
      }
      js->append("}}");
    }
    js->push_back(']');
  }
  js->push_back('}');
  return Error::Success;
}

// The borrowed buffers that follow the header on the wire, in input order (only inputs that
// carry binary data, http_client.cc:2123-2134).
void CollectTails(const std::vector<InferInput*>& inputs, std::vector<iovec>* tails, size_t* total) {
  for (InferInput* io : inputs) {
    if (io->IsSharedMemory() || !io->BinaryData()) continue;
    io->PrepareForRequest();
    bool end_of_input = false;
    while (!end_of_input) {
      const uint8_t* buf = nullptr;
      size_t n = 0;
      io->GetNext(&buf, &n, &end_of_input);
      if (buf != nullptr && n > 0) {
        tails->push_back(iovec{const_cast<uint8_t*>(buf), n});
        *total += n;
      }
    }
  }
}

std::string Base64(const uint8_t* p, size_t n) {
  static const char* tbl = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string out;
  out.reserve((n + 2) / 3 * 4);
  for (size_t i = 0; i < n; i += 3) {
    const uint32_t a = p[i], b = i + 1 < n ? p[i + 1] : 0, c = i + 2 < n ? p[i + 2] : 0;
    const uint32_t v = (a << 16) | (b << 8) | c;
    out.push_back(tbl[(v >> 18) & 63]);
    out.push_back(tbl[(v >> 12) & 63]);
    out.push_back(i + 1 < n ? tbl[(v >> 6) & 63] : '=');
    out.push_back(i + 2 < n ? tbl[v & 63] : '=');
  }
  return out;
}

std::string UrlEncode(const std::string& s) {
  static const char* hex = "0123456789ABCDEF";
  std::string out;
  for (unsigned char c : s) {
    if (isalnum(c) || c == '-' || c == '_' || c == '.' || c == '~') {
      out.push_back(static_cast<char>(c));
    } else {
      out.push_back('%');
      out.push_back(hex[c >> 4]);
      out.push_back(hex[c & 15]);
    }
  }
  return out;
}

std::string QueryString(const Parameters& params) {
  std::string q;
  for (const auto& kv : params) {
    q += q.empty() ? "?" : "&";
    q += UrlEncode(kv.first) + "=" + UrlEncode(kv.second);
  }
  return q;
}

}  // namespace

// ---- transport ---------------------------------------------------------------------------

struct HttpResponse {
  long code = 0;
  std::map<std::string, std::string> headers;  // lower-cased names
  std::string body;
};

class HttpConnection {
 public:
  HttpConnection(const std::string& host, int port) : host_(host), port_(port) {}
  ~HttpConnection() { Close(); }

  // One request/response exchange.  `tails` follow `head_and_json` without being copied.
  // timeout_us == 0: wait forever.  Timestamps: SEND_START/END around the writes,
  // RECV_START at the first response byte, RECV_END after the last.
  Error Exchange(const std::string& head_and_json, const std::vector<iovec>& tails, uint64_t timeout_us,
                 HttpResponse* resp, RequestTimers* timer) {
    for (int attempt = 0; attempt < 2; ++attempt) {
      const bool fresh = fd_ < 0;
      if (fresh) {
        Error err = Connect();
        if (!err.IsOk()) return err;
      }
      if (timer) timer->CaptureTimestamp(RequestTimers::Kind::SEND_START);
      bool sent = Send(head_and_json, tails);
      if (timer) timer->CaptureTimestamp(RequestTimers::Kind::SEND_END);
      bool timed_out = false;
      if (sent && Receive(timeout_us, resp, timer, &timed_out)) return Error::Success;
      Close();
      if (timed_out) return Error("HTTP client failed (Deadline Exceeded): Timeout was reached");
      // a keep-alive connection the server dropped meanwhile: retry once on a new one
      if (fresh) break;
    }
    return Error("HTTP client failed: connection to " + host_ + ":" + std::to_string(port_) + " was lost");
  }

 private:
  Error Connect() {
    addrinfo hints{};
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    if (getaddrinfo(host_.c_str(), std::to_string(port_).c_str(), &hints, &res) != 0 || res == nullptr) {
      return Error("HTTP client failed: Couldn't resolve host name");
    }
    for (addrinfo* ai = res; ai != nullptr; ai = ai->ai_next) {
      const int fd = socket(ai->ai_family, ai->ai_socktype, ai->ai_protocol);
      if (fd < 0) continue;
      if (connect(fd, ai->ai_addr, ai->ai_addrlen) == 0) {
        int one = 1;
        setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
        fd_ = fd;
        break;
      }
      close(fd);
    }
    freeaddrinfo(res);
    if (fd_ < 0) return Error("HTTP client failed: Couldn't connect to server");
    return Error::Success;
  }

  void Close() {
    if (fd_ >= 0) close(fd_);
    fd_ = -1;
    pending_.clear();
  }

  bool Send(const std::string& head, const std::vector<iovec>& tails) {
    std::vector<iovec> iov;
    iov.reserve(tails.size() + 1);
    iov.push_back(iovec{const_cast<char*>(head.data()), head.size()});
    iov.insert(iov.end(), tails.begin(), tails.end());
    size_t idx = 0;
    while (idx < iov.size()) {
      msghdr msg{};
      msg.msg_iov = &iov[idx];
      msg.msg_iovlen = std::min<size_t>(iov.size() - idx, 64);
      ssize_t k = sendmsg(fd_, &msg, MSG_NOSIGNAL);
      if (k < 0) {
        if (errno == EINTR) continue;
        return false;
      }
      size_t left = static_cast<size_t>(k);
      while (idx < iov.size() && left >= iov[idx].iov_len) {
        left -= iov[idx].iov_len;
        ++idx;
      }
      if (idx < iov.size() && left > 0) {
        iov[idx].iov_base = static_cast<char*>(iov[idx].iov_base) + left;
        iov[idx].iov_len -= left;
      }
    }
    return true;
  }

  // more bytes into pending_; false on EOF / error / timeout
  bool Fill(uint64_t deadline_ns, bool* timed_out) {
    if (deadline_ns != 0) {
      const uint64_t now = NowNs();
      if (now >= deadline_ns) {
        *timed_out = true;
        return false;
      }
      pollfd pfd{fd_, POLLIN, 0};
      const int ms = static_cast<int>(std::min<uint64_t>((deadline_ns - now + 999999) / 1000000, 1u << 30));
      const int r = poll(&pfd, 1, ms);
      if (r == 0) {
        *timed_out = true;
        return false;
      }
      if (r < 0 && errno != EINTR) return false;
    }
    char tmp[65536];
    for (;;) {
      const ssize_t k = recv(fd_, tmp, sizeof(tmp), 0);
      if (k > 0) {
        pending_.append(tmp, static_cast<size_t>(k));
        return true;
      }
      if (k < 0 && errno == EINTR) continue;
      return false;
    }
  }

  static uint64_t NowNs() {
    return static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
  }

  bool Receive(uint64_t timeout_us, HttpResponse* resp, RequestTimers* timer, bool* timed_out) {
    const uint64_t deadline = timeout_us ? NowNs() + timeout_us * 1000 : 0;
    resp->headers.clear();
    resp->body.clear();
    bool first = true;
    size_t header_end;
    while ((header_end = pending_.find("\r\n\r\n")) == std::string::npos) {
      if (!Fill(deadline, timed_out)) return false;
      if (first && timer) timer->CaptureTimestamp(RequestTimers::Kind::RECV_START);
      first = false;
    }
    if (first && timer) timer->CaptureTimestamp(RequestTimers::Kind::RECV_START);
    // status line
    const size_t line_end = pending_.find("\r\n");
    const size_t sp = pending_.find(' ');
    if (sp == std::string::npos || sp > line_end) return false;
    resp->code = strtol(pending_.c_str() + sp + 1, nullptr, 10);
    size_t pos = line_end + 2;
    while (pos < header_end) {
      const size_t eol = pending_.find("\r\n", pos);
      const size_t colon = pending_.find(':', pos);
      if (colon != std::string::npos && colon < eol) {
        std::string name = pending_.substr(pos, colon - pos);
        std::transform(name.begin(), name.end(), name.begin(), [](unsigned char c) { return static_cast<char>(tolower(c)); });
        size_t vb = colon + 1;
        while (vb < eol && pending_[vb] == ' ') ++vb;
        resp->headers[name] = pending_.substr(vb, eol - vb);
      }
      pos = eol + 2;
    }
    pending_.erase(0, header_end + 4);
    auto te = resp->headers.find("transfer-encoding");
    if (te != resp->headers.end() && te->second.find("chunked") != std::string::npos) {
      for (;;) {
        size_t eol;
        while ((eol = pending_.find("\r\n")) == std::string::npos) {
          if (!Fill(deadline, timed_out)) return false;
        }
        const size_t chunk = strtoull(pending_.c_str(), nullptr, 16);
        pending_.erase(0, eol + 2);
        while (pending_.size() < chunk + 2) {
          if (!Fill(deadline, timed_out)) return false;
        }
        resp->body.append(pending_, 0, chunk);
        pending_.erase(0, chunk + 2);
        if (chunk == 0) break;
      }
    } else {
      size_t clen = 0;
      auto cl = resp->headers.find("content-length");
      if (cl != resp->headers.end()) clen = strtoull(cl->second.c_str(), nullptr, 10);
      while (pending_.size() < clen) {
        if (!Fill(deadline, timed_out)) return false;
      }
      resp->body.assign(pending_, 0, clen);
      pending_.erase(0, clen);
    }
    if (timer) timer->CaptureTimestamp(RequestTimers::Kind::RECV_END);
    auto conn = resp->headers.find("connection");
    if (conn != resp->headers.end() && conn->second == "close") Close();
    return true;
  }

  std::string host_;
  int port_;
  int fd_ = -1;
  std::string pending_;
};

}  // namespace detail

// ---- result ----------------------------------------------------------------------------------

namespace {

// Response body split into JSON header and binary outputs (http_client.cc:1043-1136)
class HttpResult : public InferResult {
 public:
  explicit HttpResult(const Error& err) : status_(err) {}
  HttpResult(std::string&& body, size_t header_length, long http_code, bool verbose) : body_(std::move(body)) {
    size_t offset = header_length;
    if (http_code == 499) {
      status_ = Error("Deadline Exceeded");
      return;
    }
    const size_t json_len = offset != 0 ? offset : body_.size();
    if (verbose) std::cout << "inference response: " << body_.substr(0, json_len) << std::endl;
    std::string perr;
    if (!Value::Parse(body_.data(), json_len, &json_, &perr)) {
      status_ = Error(perr);
      return;
    }
    if (http_code != 200) {
      const Value* e = json_.Find("error");
      status_ = (e != nullptr && e->is_string()) ? Error(e->str()) : Error("inference failed with unknown error");
      return;
    }
    const Value* outputs = json_.Find("outputs");
    if (outputs == nullptr || !outputs->is_array()) return;
    for (size_t i = 0; i < outputs->size(); ++i) {
      const Value& o = (*outputs)[i];
      const Value* name = o.Find("name");
      if (!o.is_object() || name == nullptr || !name->is_string()) {
        status_ = Error("attempt to access JSON non-string as string");
        return;
      }
      const Value* params = o.Find("parameters");
      const Value* data = o.Find("data");
      if (params != nullptr) {
        const Value* sz = params->Find("binary_data_size");
        uint64_t n = 0;
        if (sz != nullptr && sz->AsUInt(&n)) {
          if (offset + n > body_.size()) {
            status_ = Error("the response body is shorter than the binary outputs it announces");
            return;
          }
          buffers_[name->str()] = std::make_pair(reinterpret_cast<const uint8_t*>(body_.data()) + offset, static_cast<size_t>(n));
          offset += n;
        } else if (params->Find("shared_memory_byte_size") == nullptr && params->Find("shared_memory_region") == nullptr &&
                   data == nullptr) {
          // the reference fails here for any parameters object without binary_data_size;
          // outputs placed in shared memory and classification outputs carry one too
          status_ = Error("attempt to access non-existing object member 'binary_data_size'");
          return;
        }
      }
      if (buffers_.find(name->str()) == buffers_.end() && data != nullptr) {
        const Value* dt = o.Find("datatype");
        if (dt == nullptr || !dt->is_string()) {
          status_ = Error("attempt to access non-existing object member 'datatype'");
          return;
        }
        converted_.emplace_back();
        status_ = detail::JsonOutputToBinary(*data, dt->str(), &converted_.back());
        if (!status_.IsOk()) return;
        buffers_[name->str()] =
            std::make_pair(reinterpret_cast<const uint8_t*>(converted_.back().data()), converted_.back().size());
      }
      outputs_[name->str()] = &o;
    }
  }

  Error RequestStatus() const override { return status_; }
  Error ModelName(std::string* name) const override { return Member("model_name", "model name", name); }
  Error ModelVersion(std::string* version) const override { return Member("model_version", "model version", version); }
  Error Id(std::string* id) const override { return Member("id", "model id", id); }

  Error Shape(const std::string& output_name, std::vector<int64_t>* shape) const override {
    if (!status_.IsOk()) return status_;
    shape->clear();
    auto it = outputs_.find(output_name);
    if (it == outputs_.end()) return Error("The response does not contain results for output name " + output_name);
    const Value* s = it->second->Find("shape");
    if (s == nullptr) return Error("The response does not contain shape for output name " + output_name);
    for (size_t i = 0; i < s->size(); ++i) {
      int64_t d = 0;
      if (!(*s)[i].AsInt(&d)) return Error("attempt to access JSON non-signed-integer as signed-integer");
      shape->push_back(d);
    }
    return Error::Success;
  }

  Error Datatype(const std::string& output_name, std::string* datatype) const override {
    if (!status_.IsOk()) return status_;
    auto it = outputs_.find(output_name);
    if (it == outputs_.end()) return Error("The response does not contain results for output name " + output_name);
    const Value* d = it->second->Find("datatype");
    if (d == nullptr || !d->is_string()) return Error("The response does not contain datatype for output name " + output_name);
    *datatype = d->str();
    return Error::Success;
  }

  Error RawData(const std::string& output_name, const uint8_t** buf, size_t* byte_size) const override {
    if (!status_.IsOk()) return status_;
    auto it = buffers_.find(output_name);
    if (it == buffers_.end()) return Error("The response does not contain results for output name " + output_name);
    *buf = it->second.first;
    *byte_size = it->second.second;
    return Error::Success;
  }

  Error IsFinalResponse(bool* is_final_response) const override {
    if (is_final_response == nullptr) return Error("is_final_response cannot be nullptr");
    *is_final_response = true;
    return Error::Success;
  }

  Error IsNullResponse(bool* is_null_response) const override {
    if (is_null_response == nullptr) return Error("is_null_response cannot be nullptr");
    *is_null_response = false;
    return Error::Success;
  }

  Error StringData(const std::string& output_name, std::vector<std::string>* string_result) const override {
    if (!status_.IsOk()) return status_;
    std::string datatype;
    Error err = Datatype(output_name, &datatype);
    if (!err.IsOk()) return err;
    if (datatype != "BYTES") {
      return Error("This function supports tensors with datatype 'BYTES', requested output tensor '" + output_name +
                   "' with datatype '" + datatype + "'");
    }
    const uint8_t* buf = nullptr;
    size_t byte_size = 0;
    err = RawData(output_name, &buf, &byte_size);
    if (!err.IsOk()) return err;
    string_result->clear();
    size_t pos = 0;
    while (pos + sizeof(uint32_t) <= byte_size) {
      uint32_t len;
      memcpy(&len, buf + pos, sizeof(len));
      if (pos + sizeof(len) + len > byte_size) return Error("malformed BYTES tensor in the response");
      string_result->emplace_back(reinterpret_cast<const char*>(buf + pos + sizeof(len)), len);
      pos += sizeof(len) + len;
    }
    return Error::Success;
  }

  std::string DebugString() const override { return status_.IsOk() ? json_.Dump() : status_.Message(); }

 private:
  Error Member(const char* key, const char* what, std::string* out) const {
    if (!status_.IsOk()) return status_;
    const Value* v = json_.Find(key);
    if (v == nullptr || !v->is_string()) return Error(std::string(what) + " was not returned in the response");
    *out = v->str();
    return Error::Success;
  }

  Error status_;
  std::string body_;
  Value json_;
  std::map<std::string, const Value*> outputs_;
  std::map<std::string, std::pair<const uint8_t*, size_t>> buffers_;
  std::deque<std::string> converted_;
};

}  // namespace

// ---- client -------------------------------------------------------------------------------------

namespace detail {

// zlib ("deflate") / gzip stream of `body`, produced on the device (tb200_deflate_async): the
// request side of http_client.cc:146-221.  There is no host encoder in this library.
Error DeflateOnDevice(std::string* body, bool gzip) {
  int count = 0;
  if (tb200_device_count(&count) != TB200_OK || count < 1) {
    return Error("request compression runs on the device (tb200_deflate_async) and no CUDA device is available");
  }
  tb200_ctx* ctx = nullptr;
  if (tb200_ctx_create(0, &ctx) != TB200_OK) return Error(std::string("request compression: ") + tb200_last_error());
  const uint64_t n = body->size();
  const uint64_t cap = tb200_deflate_bound(n);
  void *src = nullptr, *dst = nullptr, *host = nullptr, *host_dev = nullptr;
  Error err;
  auto check = [&err](int rc) {
    if (rc != TB200_OK && err.IsOk()) err = Error(std::string("request compression: ") + tb200_last_error());
    return rc == TB200_OK;
  };
  const uint64_t size_off = (cap + 15) & ~static_cast<uint64_t>(15);  // pinned staging: [stream | pad | uint64 size]
  if (check(tb200_device_alloc(0, n + 16, &src)) && check(tb200_device_alloc(0, cap + 16, &dst)) &&
      check(tb200_host_alloc(size_off + 16, &host, &host_dev)) && check(tb200_memcpy_h2d_async(ctx, src, body->data(), n)) &&
      check(tb200_deflate_async(ctx, dst, cap, src, n, gzip ? TB200_DEFLATE_GZIP : TB200_DEFLATE_ZLIB,
                                reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(host_dev) + size_off))) &&
      check(tb200_ctx_sync(ctx))) {
    uint64_t out_size = 0;
    memcpy(&out_size, static_cast<uint8_t*>(host) + size_off, 8);
    if (out_size == 0 || out_size > cap) {
      err = Error("request compression: the device encoder reported an invalid stream size");
    } else if (check(tb200_memcpy_d2h_async(ctx, host, dst, out_size)) && check(tb200_ctx_sync(ctx))) {
      body->assign(static_cast<const char*>(host), out_size);
    }
  }
  if (src) tb200_device_free(0, src);
  if (dst) tb200_device_free(0, dst);
  if (host) tb200_host_free(host);
  tb200_ctx_destroy(ctx);
  return err;
}

// zlib or gzip stream -> bytes (response side, http_client.cc:2211-2254: zlib on the host)
Error Inflate(const std::string& in, std::string* out) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 15 + 32) != Z_OK) return Error("failed to initialise zlib");  // +32: zlib or gzip header
  zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data()));
  zs.avail_in = static_cast<uInt>(in.size());
  out->clear();
  char buf[65536];
  int rc = Z_OK;
  while (rc == Z_OK) {
    zs.next_out = reinterpret_cast<Bytef*>(buf);
    zs.avail_out = sizeof(buf);
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc == Z_OK || rc == Z_STREAM_END) out->append(buf, sizeof(buf) - zs.avail_out);
    if (rc == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) break;  // truncated stream
  }
  inflateEnd(&zs);
  if (rc != Z_STREAM_END) return Error("failed to decompress the response body");
  return Error::Success;
}

}  // namespace detail

struct InferenceServerHttpClient::AsyncJob {
  CompressionType request_compression = CompressionType::NONE, response_compression = CompressionType::NONE;
  OnCompleteFn callback;
  OnMultiCompleteFn multi_callback;
  std::vector<InferOptions> options;
  std::vector<std::vector<InferInput*>> inputs;
  std::vector<std::vector<const InferRequestedOutput*>> outputs;
  Headers headers;
  Parameters query_params;
};

InferenceServerHttpClient::InferenceServerHttpClient(const std::string& host, int port, const std::string& base_path, bool verbose)
    : InferenceServerClient(verbose), host_(host), port_(port), base_path_(base_path),
      sync_conn_(new detail::HttpConnection(host, port)) {}

InferenceServerHttpClient::~InferenceServerHttpClient() {
  {
    std::lock_guard<std::mutex> lk(async_mu_);
    exiting_ = true;
  }
  async_cv_.notify_all();
  if (worker_.joinable()) worker_.join();
}

Error InferenceServerHttpClient::Create(std::unique_ptr<InferenceServerHttpClient>* client, const std::string& server_url,
                                        bool verbose, const HttpSslOptions&) {
  std::string url = server_url;
  if (url.rfind("https://", 0) == 0) return Error("https is not supported: this build has no TLS transport");
  if (url.rfind("http://", 0) == 0) url.erase(0, 7);
  std::string base_path;
  const size_t slash = url.find('/');
  if (slash != std::string::npos) {
    base_path = url.substr(slash);
    while (!base_path.empty() && base_path.back() == '/') base_path.pop_back();
    url.resize(slash);
  }
  int port = 80;
  const size_t colon = url.rfind(':');
  if (colon != std::string::npos) {
    port = atoi(url.c_str() + colon + 1);
    url.resize(colon);
  }
  if (url.empty() || port <= 0) return Error("failed to parse the server url '" + server_url + "'");
  client->reset(new InferenceServerHttpClient(url, port, base_path, verbose));
  return Error::Success;
}

namespace {
std::string RequestHead(const char* method, const std::string& host, int port, const std::string& target, const Headers& headers,
                        const std::vector<std::pair<std::string, std::string>>& extra, size_t content_length) {
  std::string head = std::string(method) + " " + target + " HTTP/1.1\r\nHost: " + host + ":" + std::to_string(port) +
                     "\r\nUser-Agent: tb200-client/1.0\r\nAccept: */*\r\n";
  for (const auto& kv : extra) head += kv.first + ": " + kv.second + "\r\n";
  for (const auto& kv : headers) head += kv.first + ": " + kv.second + "\r\n";
  if (strcmp(method, "POST") == 0) head += "Content-Length: " + std::to_string(content_length) + "\r\n";
  head += "\r\n";
  return head;
}

Error StatusError(const detail::HttpResponse& resp) {
  // error responses carry {"error": "..."} (http_client.cc:2257-2279)
  Value js;
  std::string perr;
  if (!resp.body.empty() && Value::Parse(resp.body.data(), resp.body.size(), &js, &perr)) {
    const Value* e = js.Find("error");
    if (e != nullptr && e->is_string()) return Error(e->str());
  }
  return Error("[HTTP " + std::to_string(resp.code) + "] " + (resp.body.empty() ? std::string("request failed") : resp.body));
}
}  // namespace

Error InferenceServerHttpClient::Get(const std::string& path, const Headers& headers, const Parameters& query_params,
                                     std::string* response, long* http_code) {
  std::lock_guard<std::mutex> lk(sync_mu_);
  const std::string head = RequestHead("GET", host_, port_, base_path_ + path + detail::QueryString(query_params), headers, {}, 0);
  if (verbose_) std::cout << "GET " << path << std::endl;
  detail::HttpResponse resp;
  Error err = sync_conn_->Exchange(head, {}, 0, &resp, nullptr);
  if (!err.IsOk()) return err;
  if (http_code != nullptr) *http_code = resp.code;
  if (response != nullptr) *response = resp.body;
  if (verbose_) std::cout << resp.body << std::endl;
  if (resp.code != 200 && http_code == nullptr) return StatusError(resp);
  return Error::Success;
}

Error InferenceServerHttpClient::Post(const std::string& path, const std::string& request, const Headers& headers,
                                      const Parameters& query_params, std::string* response, long* http_code) {
  std::lock_guard<std::mutex> lk(sync_mu_);
  std::string head = RequestHead("POST", host_, port_, base_path_ + path + detail::QueryString(query_params), headers,
                                 {{"Content-Type", "application/json"}}, request.size());
  if (verbose_) std::cout << "POST " << path << ", body " << request << std::endl;
  head += request;
  detail::HttpResponse resp;
  Error err = sync_conn_->Exchange(head, {}, 0, &resp, nullptr);
  if (!err.IsOk()) return err;
  if (http_code != nullptr) *http_code = resp.code;
  if (response != nullptr) *response = resp.body;
  if (verbose_) std::cout << resp.body << std::endl;
  if (resp.code != 200) return StatusError(resp);
  return Error::Success;
}

Error InferenceServerHttpClient::IsServerLive(bool* live, const Headers& headers, const Parameters& query_params) {
  long code = 0;
  Error err = Get("/v2/health/live", headers, query_params, nullptr, &code);
  *live = code == 200;
  return err;
}

Error InferenceServerHttpClient::IsServerReady(bool* ready, const Headers& headers, const Parameters& query_params) {
  long code = 0;
  Error err = Get("/v2/health/ready", headers, query_params, nullptr, &code);
  *ready = code == 200;
  return err;
}

namespace {
std::string ModelPath(const std::string& model_name, const std::string& model_version) {
  std::string p = "/v2/models/" + model_name;
  if (!model_version.empty()) p += "/versions/" + model_version;
  return p;
}
}  // namespace

Error InferenceServerHttpClient::IsModelReady(bool* ready, const std::string& model_name, const std::string& model_version,
                                              const Headers& headers, const Parameters& query_params) {
  long code = 0;
  Error err = Get(ModelPath(model_name, model_version) + "/ready", headers, query_params, nullptr, &code);
  *ready = code == 200;
  return err;
}

Error InferenceServerHttpClient::ServerMetadata(std::string* server_metadata, const Headers& headers, const Parameters& query_params) {
  return Get("/v2", headers, query_params, server_metadata);
}

Error InferenceServerHttpClient::ModelMetadata(std::string* model_metadata, const std::string& model_name,
                                               const std::string& model_version, const Headers& headers,
                                               const Parameters& query_params) {
  return Get(ModelPath(model_name, model_version), headers, query_params, model_metadata);
}

Error InferenceServerHttpClient::ModelConfig(std::string* model_config, const std::string& model_name,
                                             const std::string& model_version, const Headers& headers,
                                             const Parameters& query_params) {
  return Get(ModelPath(model_name, model_version) + "/config", headers, query_params, model_config);
}

Error InferenceServerHttpClient::ModelRepositoryIndex(std::string* repository_index, const Headers& headers,
                                                      const Parameters& query_params) {
  return Post("/v2/repository/index", "", headers, query_params, repository_index);
}

Error InferenceServerHttpClient::LoadModel(const std::string& model_name, const Headers& headers, const Parameters& query_params,
                                           const std::string& config, const std::map<std::string, std::vector<char>>& files) {
  // {"parameters":{"config":..., "file:<path>": base64}} (http_client.cc:1462-1504)
  std::string body;
  if (!config.empty() || !files.empty()) {
    Value params = Value::Object();
    if (!config.empty()) params.Add("config", Value::String(config));
    for (const auto& kv : files) {
      params.Add(kv.first, Value::String(detail::Base64(reinterpret_cast<const uint8_t*>(kv.second.data()), kv.second.size())));
    }
    Value req = Value::Object();
    req.Add("parameters", std::move(params));
    body = req.Dump();
  }
  std::string response;
  return Post("/v2/repository/models/" + model_name + "/load", body, headers, query_params, &response);
}

Error InferenceServerHttpClient::UnloadModel(const std::string& model_name, const Headers& headers, const Parameters& query_params) {
  std::string response;
  return Post("/v2/repository/models/" + model_name + "/unload", "", headers, query_params, &response);
}

Error InferenceServerHttpClient::ModelInferenceStatistics(std::string* infer_stat, const std::string& model_name,
                                                          const std::string& model_version, const Headers& headers,
                                                          const Parameters& query_params) {
  std::string p = "/v2/models";
  if (!model_name.empty()) p += "/" + model_name;
  if (!model_version.empty()) p += "/versions/" + model_version;
  return Get(p + "/stats", headers, query_params, infer_stat);
}

Error InferenceServerHttpClient::UpdateTraceSettings(std::string* response, const std::string& model_name,
                                                     const std::map<std::string, std::vector<std::string>>& settings,
                                                     const Headers& headers, const Parameters& query_params) {
  // empty value list clears a setting (null), one value is a scalar, several an array
  Value req = Value::Object();
  for (const auto& kv : settings) {
    if (kv.second.empty()) {
      req.Add(kv.first, Value());
    } else if (kv.second.size() == 1) {
      req.Add(kv.first, Value::String(kv.second[0]));
    } else {
      Value arr = Value::Array();
      for (const std::string& s : kv.second) arr.Append(Value::String(s));
      req.Add(kv.first, std::move(arr));
    }
  }
  const std::string p = model_name.empty() ? "/v2/trace/setting" : "/v2/models/" + model_name + "/trace/setting";
  return Post(p, req.Dump(), headers, query_params, response);
}

Error InferenceServerHttpClient::GetTraceSettings(std::string* settings, const std::string& model_name, const Headers& headers,
                                                  const Parameters& query_params) {
  const std::string p = model_name.empty() ? "/v2/trace/setting" : "/v2/models/" + model_name + "/trace/setting";
  return Get(p, headers, query_params, settings);
}

Error InferenceServerHttpClient::SystemSharedMemoryStatus(std::string* status, const std::string& region_name,
                                                          const Headers& headers, const Parameters& query_params) {
  std::string p = "/v2/systemsharedmemory";
  if (!region_name.empty()) p += "/region/" + region_name;
  return Get(p + "/status", headers, query_params, status);
}

Error InferenceServerHttpClient::RegisterSystemSharedMemory(const std::string& name, const std::string& key, const size_t byte_size,
                                                            const size_t offset, const Headers& headers,
                                                            const Parameters& query_params) {
  Value req = Value::Object();
  req.Add("key", Value::String(key));
  req.Add("offset", Value::UInt(offset));
  req.Add("byte_size", Value::UInt(byte_size));
  std::string response;
  return Post("/v2/systemsharedmemory/region/" + name + "/register", req.Dump(), headers, query_params, &response);
}

Error InferenceServerHttpClient::UnregisterSystemSharedMemory(const std::string& name, const Headers& headers,
                                                              const Parameters& query_params) {
  std::string p = "/v2/systemsharedmemory";
  if (!name.empty()) p += "/region/" + name;
  std::string response;
  return Post(p + "/unregister", "", headers, query_params, &response);
}

Error InferenceServerHttpClient::CudaSharedMemoryStatus(std::string* status, const std::string& region_name,
                                                        const Headers& headers, const Parameters& query_params) {
  std::string p = "/v2/cudasharedmemory";
  if (!region_name.empty()) p += "/region/" + region_name;
  return Get(p + "/status", headers, query_params, status);
}

Error InferenceServerHttpClient::RegisterCudaSharedMemoryRaw(const std::string& name, const uint8_t* handle64, const size_t device_id,
                                                             const size_t byte_size, const Headers& headers,
                                                             const Parameters& query_params) {
  // {"raw_handle":{"b64":...},"device_id":..,"byte_size":..} (http_client.cc:1706-1748)
  Value raw = Value::Object();
  raw.Add("b64", Value::String(detail::Base64(handle64, 64)));
  Value req = Value::Object();
  req.Add("raw_handle", std::move(raw));
  req.Add("device_id", Value::UInt(device_id));
  req.Add("byte_size", Value::UInt(byte_size));
  std::string response;
  return Post("/v2/cudasharedmemory/region/" + name + "/register", req.Dump(), headers, query_params, &response);
}

Error InferenceServerHttpClient::UnregisterCudaSharedMemory(const std::string& name, const Headers& headers,
                                                            const Parameters& query_params) {
  std::string p = "/v2/cudasharedmemory";
  if (!name.empty()) p += "/region/" + name;
  std::string response;
  return Post(p + "/unregister", "", headers, query_params, &response);
}

Error InferenceServerHttpClient::GenerateRequestBody(std::vector<char>* request_body, size_t* header_length,
                                                     const InferOptions& options, const std::vector<InferInput*>& inputs,
                                                     const std::vector<const InferRequestedOutput*>& outputs) {
  std::string js;
  Error err = detail::BuildInferHeader(options, inputs, outputs, &js);
  if (!err.IsOk()) return err;
  std::vector<iovec> tails;
  size_t tail_bytes = 0;
  detail::CollectTails(inputs, &tails, &tail_bytes);
  *header_length = js.size();
  request_body->resize(js.size() + tail_bytes);
  memcpy(request_body->data(), js.data(), js.size());
  size_t pos = js.size();
  for (const iovec& v : tails) {
    memcpy(request_body->data() + pos, v.iov_base, v.iov_len);
    pos += v.iov_len;
  }
  return Error::Success;
}

Error InferenceServerHttpClient::ParseResponseBody(InferResult** result, const std::vector<char>& response_body,
                                                   size_t header_length) {
  *result = new HttpResult(std::string(response_body.begin(), response_body.end()), header_length, 200, false);
  return Error::Success;
}

Error InferenceServerHttpClient::InferOn(detail::HttpConnection* conn, InferResult** result, const InferOptions& options,
                                         const std::vector<InferInput*>& inputs,
                                         const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                         const Parameters& query_params, CompressionType request_compression,
                                         CompressionType response_compression) {
  RequestTimers timer;
  timer.CaptureTimestamp(RequestTimers::Kind::REQUEST_START);
  std::string js;
  Error err = detail::BuildInferHeader(options, inputs, outputs, &js);
  if (!err.IsOk()) return err;
  std::vector<iovec> tails;
  size_t tail_bytes = 0;
  detail::CollectTails(inputs, &tails, &tail_bytes);
  bool all_json = true;
  for (const InferInput* io : inputs) {
    if (io->BinaryData()) all_json = false;
  }
  std::vector<std::pair<std::string, std::string>> extra = {{kInferHeaderContentLengthHTTPHeader, std::to_string(js.size())},
                                                            {"Content-Type", all_json ? "application/json" : "application/octet-stream"}};
  if (response_compression != CompressionType::NONE) {
    extra.emplace_back("Accept-Encoding", response_compression == CompressionType::GZIP ? "gzip" : "deflate");
  }
  if (verbose_) std::cout << "inference request: " << js << std::endl;
  std::string payload = js;
  size_t content_length = js.size() + tail_bytes;
  if (request_compression != CompressionType::NONE) {
    // the whole body (JSON + tensors) becomes one zlib / gzip stream made on the device; the
    // Inference-Header-Content-Length keeps the uncompressed JSON size (http_client.cc:1498-1527)
    payload.reserve(content_length);
    for (const iovec& v : tails) payload.append(static_cast<const char*>(v.iov_base), v.iov_len);
    tails.clear();
    err = detail::DeflateOnDevice(&payload, request_compression == CompressionType::GZIP);
    if (!err.IsOk()) return err;
    content_length = payload.size();
    extra.emplace_back("Content-Encoding", request_compression == CompressionType::GZIP ? "gzip" : "deflate");
  }
  std::string head = RequestHead(
      "POST", host_, port_, base_path_ + ModelPath(options.model_name_, options.model_version_) + "/infer" + detail::QueryString(query_params),
      headers, extra, content_length);
  head += payload;
  detail::HttpResponse resp;
  err = conn->Exchange(head, tails, options.client_timeout_, &resp, &timer);
  if (!err.IsOk()) return err;
  const auto encoding = resp.headers.find("content-encoding");
  if (encoding != resp.headers.end() && (encoding->second == "gzip" || encoding->second == "deflate")) {
    std::string plain;
    err = detail::Inflate(resp.body, &plain);
    if (!err.IsOk()) return err;
    resp.body = std::move(plain);
  }
  size_t header_length = 0;
  auto it = resp.headers.find("inference-header-content-length");
  if (it != resp.headers.end()) header_length = strtoull(it->second.c_str(), nullptr, 10);
  *result = new HttpResult(std::move(resp.body), header_length, resp.code, verbose_);
  timer.CaptureTimestamp(RequestTimers::Kind::REQUEST_END);
  err = UpdateInferStat(timer);
  if (!err.IsOk()) std::cerr << "Failed to update context stat: " << err << std::endl;
  return (*result)->RequestStatus();
}

Error InferenceServerHttpClient::Infer(InferResult** result, const InferOptions& options, const std::vector<InferInput*>& inputs,
                                       const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                       const Parameters& query_params, const CompressionType request_compression_algorithm,
                                       const CompressionType response_compression_algorithm) {
  std::lock_guard<std::mutex> lk(sync_mu_);
  return InferOn(sync_conn_.get(), result, options, inputs, outputs, headers, query_params, request_compression_algorithm,
                 response_compression_algorithm);
}

void InferenceServerHttpClient::AsyncWorker() {
  detail::HttpConnection conn(host_, port_);
  for (;;) {
    std::shared_ptr<AsyncJob> job;
    {
      std::unique_lock<std::mutex> lk(async_mu_);
      async_cv_.wait(lk, [&] { return exiting_ || !async_jobs_.empty(); });
      if (async_jobs_.empty()) return;
      job = async_jobs_.front();
      async_jobs_.pop_front();
    }
    std::vector<InferResult*> results;
    for (size_t i = 0; i < job->options.size(); ++i) {
      InferResult* r = nullptr;
      static const std::vector<const InferRequestedOutput*> none;
      const auto& outs = job->outputs.empty() ? none : job->outputs[job->outputs.size() == 1 ? 0 : i];
      Error err = InferOn(&conn, &r, job->options[i], job->inputs[i], outs, job->headers, job->query_params, job->request_compression,
                          job->response_compression);
      if (r == nullptr) r = new HttpResult(err);
      results.push_back(r);
    }
    if (job->callback) job->callback(results[0]);
    else job->multi_callback(results);
  }
}

Error InferenceServerHttpClient::AsyncInfer(OnCompleteFn callback, const InferOptions& options, const std::vector<InferInput*>& inputs,
                                            const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                            const Parameters& query_params, const CompressionType request_compression_algorithm,
                                            const CompressionType response_compression_algorithm) {
  if (callback == nullptr) return Error("Callback function must be provided along with AsyncInfer() call.");
  auto job = std::make_shared<AsyncJob>();
  job->request_compression = request_compression_algorithm;
  job->response_compression = response_compression_algorithm;
  job->callback = std::move(callback);
  job->options.push_back(options);
  job->inputs.push_back(inputs);
  job->outputs.push_back(outputs);
  job->headers = headers;
  job->query_params = query_params;
  {
    std::lock_guard<std::mutex> lk(async_mu_);
    if (!worker_.joinable()) worker_ = std::thread(&InferenceServerHttpClient::AsyncWorker, this);
    async_jobs_.push_back(std::move(job));
  }
  async_cv_.notify_one();
  return Error::Success;
}

namespace {
Error CheckMulti(const std::vector<InferOptions>& options, const std::vector<std::vector<InferInput*>>& inputs,
                 const std::vector<std::vector<const InferRequestedOutput*>>& outputs) {
  // http_client.cc:1907-1925
  if (options.size() != 1 && options.size() != inputs.size()) {
    return Error("'options' must either contain 1 element or match size of 'inputs'");
  }
  if (outputs.size() > 1 && outputs.size() != inputs.size()) {
    return Error("'outputs' must either contain 0/1 element or match size of 'inputs'");
  }
  return Error::Success;
}
}  // namespace

Error InferenceServerHttpClient::InferMulti(std::vector<InferResult*>* results, const std::vector<InferOptions>& options,
                                            const std::vector<std::vector<InferInput*>>& inputs,
                                            const std::vector<std::vector<const InferRequestedOutput*>>& outputs,
                                            const Headers& headers, const Parameters& query_params,
                                            const CompressionType request_compression_algorithm,
                                            const CompressionType response_compression_algorithm) {
  Error err = CheckMulti(options, inputs, outputs);
  if (!err.IsOk()) return err;
  static const std::vector<const InferRequestedOutput*> none;
  for (size_t i = 0; i < inputs.size(); ++i) {
    const InferOptions& opt = options.size() == 1 ? options[0] : options[i];
    const auto& outs = outputs.empty() ? none : (outputs.size() == 1 ? outputs[0] : outputs[i]);
    results->emplace_back();
    err = Infer(&results->back(), opt, inputs[i], outs, headers, query_params, request_compression_algorithm,
                response_compression_algorithm);
    if (!err.IsOk()) return err;
  }
  return Error::Success;
}

Error InferenceServerHttpClient::AsyncInferMulti(OnMultiCompleteFn callback, const std::vector<InferOptions>& options,
                                                 const std::vector<std::vector<InferInput*>>& inputs,
                                                 const std::vector<std::vector<const InferRequestedOutput*>>& outputs,
                                                 const Headers& headers, const Parameters& query_params,
                                                 const CompressionType request_compression_algorithm,
                                                 const CompressionType response_compression_algorithm) {
  if (callback == nullptr) return Error("Callback function must be provided along with AsyncInferMulti() call.");
  Error err = CheckMulti(options, inputs, outputs);
  if (!err.IsOk()) return err;
  auto job = std::make_shared<AsyncJob>();
  job->request_compression = request_compression_algorithm;
  job->response_compression = response_compression_algorithm;
  job->multi_callback = std::move(callback);
  for (size_t i = 0; i < inputs.size(); ++i) job->options.push_back(options.size() == 1 ? options[0] : options[i]);
  job->inputs = inputs;
  job->outputs = outputs;
  job->headers = headers;
  job->query_params = query_params;
  {
    std::lock_guard<std::mutex> lk(async_mu_);
    if (!worker_.joinable()) worker_ = std::thread(&InferenceServerHttpClient::AsyncWorker, this);
    async_jobs_.push_back(std::move(job));
  }
  async_cv_.notify_one();
  return Error::Success;
}

// ---- CudaRegion: device memory through the C ABI ---------------------------------------------------

namespace {
Error Native(int rc, const char* what) {
  if (rc == TB200_OK) return Error::Success;
  const char* msg = tb200_last_error();
  return Error(std::string(what) + ": " + (msg != nullptr ? msg : "unknown error"));
}
}  // namespace

CudaRegion::~CudaRegion() {
  if (region_ != nullptr) tb200_region_destroy(region_);
  if (ctx_ != nullptr) tb200_ctx_destroy(ctx_);
}

Error CudaRegion::Create(std::unique_ptr<CudaRegion>* region, const std::string& name, size_t byte_size, int device_id) {
  std::unique_ptr<CudaRegion> r(new CudaRegion());
  r->name_ = name;
  r->byte_size_ = byte_size;
  r->device_id_ = device_id;
  Error err = Native(tb200_ctx_create(device_id, &r->ctx_), "unable to create a device context");
  if (!err.IsOk()) return err;
  err = Native(tb200_region_create(name.c_str(), byte_size, device_id, &r->region_), "unable to create cuda shared memory handle");
  if (!err.IsOk()) return err;
  *region = std::move(r);
  return Error::Success;
}

void* CudaRegion::DevicePtr() const { return reinterpret_cast<void*>(tb200_region_base(region_)); }

Error CudaRegion::IpcHandle(uint8_t out[64]) const {
  return Native(tb200_region_ipc_handle(region_, out), "unable to export the CUDA IPC handle");
}

Error CudaRegion::Register(InferenceServerHttpClient* client) const {
  uint8_t handle[64];
  Error err = IpcHandle(handle);
  if (!err.IsOk()) return err;
  return client->RegisterCudaSharedMemoryRaw(name_, handle, static_cast<size_t>(device_id_), byte_size_);
}

Error CudaRegion::SetFromInput(InferInput& input, size_t offset) {
  std::vector<const void*> srcs;
  std::vector<uint64_t> sizes;
  input.PrepareForRequest();
  bool end_of_input = false;
  while (!end_of_input) {
    const uint8_t* buf = nullptr;
    size_t n = 0;
    input.GetNext(&buf, &n, &end_of_input);
    if (buf != nullptr && n > 0) {
      srcs.push_back(buf);
      sizes.push_back(n);
    }
  }
  return Native(tb200_region_write_host_gather(ctx_, region_, offset, static_cast<int>(srcs.size()), srcs.data(), sizes.data()),
                "unable to set values in cuda shared memory");
}

Error CudaRegion::Write(size_t offset, const void* src, size_t byte_size) {
  return Native(tb200_region_write_host(ctx_, region_, offset, src, byte_size), "unable to set values in cuda shared memory");
}

Error CudaRegion::Read(size_t offset, void* dst, size_t byte_size) const {
  return Native(tb200_region_read_host(ctx_, region_, offset, dst, byte_size), "failed to read cuda shared memory results");
}

Error CudaRegion::FillRandom(size_t offset, const std::string& datatype, size_t byte_size, uint64_t seed, uint64_t stream_id,
                             bool zero) {
  const uint32_t dt = tb200_dtype_from_name(datatype.c_str());
  if (dt == TB200_INVALID || dt == TB200_BYTES) return Error("datatype '" + datatype + "' cannot be generated on the device");
  if (offset + byte_size > byte_size_) return Error("the tensor does not fit into the region");
  tb200_fill_job job{};
  job.dst = tb200_region_base(region_) + offset;
  job.nbytes = byte_size;
  job.stream = stream_id;
  job.dtype = dt;
  job.mode = zero ? TB200_FILL_ZERO : TB200_FILL_RANDOM;
  Error err = Native(tb200_fill_async(ctx_, &job, 1, seed, 0), "device fill failed");
  if (!err.IsOk()) return err;
  return Native(tb200_ctx_sync(ctx_), "device fill failed");
}

Error CudaRegion::CheckAddSub(size_t out0_offset, size_t out1_offset, const CudaRegion& inputs, size_t in0_offset,
                              size_t in1_offset, size_t byte_size, uint64_t* mismatches) const {
  void *host = nullptr, *dev = nullptr;
  Error err = Native(tb200_host_alloc(sizeof(tb200_check_result), &host, &dev), "unable to allocate the result slot");
  if (!err.IsOk()) return err;
  tb200_check_job job{};
  job.a = tb200_region_base(region_) + out0_offset;
  job.b = tb200_region_base(region_) + out1_offset;
  job.c = tb200_region_base(inputs.region_) + in0_offset;
  job.d = tb200_region_base(inputs.region_) + in1_offset;
  job.nbytes = byte_size;
  job.kind = TB200_CHECK_ADDSUB;
  err = Native(tb200_check_async(ctx_, &job, 1, static_cast<tb200_check_result*>(dev)), "device check failed");
  if (err.IsOk()) err = Native(tb200_ctx_sync(ctx_), "device check failed");
  if (err.IsOk()) *mismatches = static_cast<const tb200_check_result*>(host)->mismatches;
  tb200_host_free(host);
  return err;
}

}}  // namespace tb200::client
