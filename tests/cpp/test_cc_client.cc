// Known-answer and loopback tests of the C++ front end (client_b200/cpp).
//
// The conversion cases restate the reference's HTTPJSONDataTest
// (src/c++/tests/cc_client_test.cc:1662-2170) value for value; the InferMulti / AsyncInfer
// cases follow its ClientTest fixture (:300-1200) against the `simple` add/sub model.
// Usage: test_cc_client [host:port]   (without a URL only the offline cases run)
#include <array>
#include <atomic>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <iostream>

#include "common.h"
#include "http_client.h"
#include "json.h"

namespace tc = triton::client;

static int g_failures = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      ++g_failures;                                                              \
      std::cerr << __FILE__ << ":" << __LINE__ << ": CHECK failed: " #cond "\n"; \
    }                                                                            \
  } while (0)
#define CHECK_OK(err)                                                                               \
  do {                                                                                              \
    const tc::Error e__ = (err);                                                                    \
    if (!e__.IsOk()) {                                                                              \
      ++g_failures;                                                                                 \
      std::cerr << __FILE__ << ":" << __LINE__ << ": unexpected error: " << e__.Message() << "\n"; \
    }                                                                                               \
  } while (0)

template <typename T>
static const uint8_t* Bytes(const T& a) {
  return reinterpret_cast<const uint8_t*>(a.data());
}

static std::vector<std::string> Items(const uint8_t* buf, size_t n, const std::string& dt, bool expect_ok = true) {
  std::vector<std::string> items;
  tc::Error err = tc::detail::BinaryInputToJsonText(buf, n, dt, &items);
  CHECK(err.IsOk() == expect_ok);
  return items;
}

static void TestBinaryInputsToJson() {
  // two AppendRaw buffers of a [1,2,2] INT32 tensor flatten into one 8-element array
  tc::InferInput* input = nullptr;
  tc::InferInput::Create(&input, "INPUT", {1, 2, 2}, "INT32");
  int32_t raw1[4] = {1, 3, 5, 7};
  int32_t raw2[4] = {2, 4, 6, 8};
  input->AppendRaw(reinterpret_cast<uint8_t*>(raw1), sizeof(raw1));
  input->AppendRaw(reinterpret_cast<uint8_t*>(raw2), sizeof(raw2));
  std::vector<std::string> items;
  CHECK_OK(tc::detail::BinaryInputsToJsonText(*input, &items));
  const std::vector<std::string> want = {"1", "3", "5", "7", "2", "4", "6", "8"};
  CHECK(items == want);
  size_t byte_size = 0;
  input->ByteSize(&byte_size);
  CHECK(byte_size == 32);
  delete input;
}

static void TestBinaryInputToJson() {
  using V = std::vector<std::string>;
  std::array<bool, 2> b({false, true});
  CHECK((Items(Bytes(b), 2, "BOOL") == V{"false", "true"}));
  std::array<uint8_t, 2> u8({1, UINT8_MAX});
  CHECK((Items(Bytes(u8), 2, "UINT8") == V{"1", "255"}));
  std::array<uint16_t, 2> u16({1, UINT16_MAX});
  CHECK((Items(Bytes(u16), 2, "UINT16") == V{"1", "65535"}));
  std::array<uint32_t, 2> u32({1, UINT32_MAX});
  CHECK((Items(Bytes(u32), 2, "UINT32") == V{"1", "4294967295"}));
  std::array<uint64_t, 2> u64({1, UINT64_MAX});
  CHECK((Items(Bytes(u64), 2, "UINT64") == V{"1", "18446744073709551615"}));
  std::array<int8_t, 2> i8({INT8_MIN, INT8_MAX});
  CHECK((Items(Bytes(i8), 2, "INT8") == V{"-128", "127"}));
  std::array<int16_t, 2> i16({INT16_MIN, INT16_MAX});
  CHECK((Items(Bytes(i16), 2, "INT16") == V{"-32768", "32767"}));
  std::array<int32_t, 2> i32({INT32_MIN, INT32_MAX});
  CHECK((Items(Bytes(i32), 2, "INT32") == V{"-2147483648", "2147483647"}));
  std::array<int64_t, 2> i64({INT64_MIN, INT64_MAX});
  CHECK((Items(Bytes(i64), 2, "INT64") == V{"-9223372036854775808", "9223372036854775807"}));
  Items(Bytes(u16), 2, "FP16", false);
  std::array<float, 2> f32({-1000.0f, 1000.0f});
  CHECK((Items(Bytes(f32), 2, "FP32") == V{"-1000.0", "1000.0"}));
  std::array<double, 2> f64({-1000.0, 1000.0});
  CHECK((Items(Bytes(f64), 2, "FP64") == V{"-1000.0", "1000.0"}));
  std::array<uint8_t, 12> bytes({2, 0, 0, 0, 1, INT8_MAX, 2, 0, 0, 0, 2, INT8_MAX});
  const V s = Items(bytes.data(), 2, "BYTES");
  CHECK(s.size() == 2 && s[0] == "\"\\u0001\x7f\"" && s[1] == "\"\\u0002\x7f\"");
  Items(Bytes(u16), 2, "BF16", false);
  Items(Bytes(u16), 2, "invaliddatatype", false);
}

static void TestDoubleFormatting() {
  auto fmt = [](double v) {
    std::string s;
    tb200::json::Value::WriteDouble(v, &s);
    return s;
  };
  CHECK(fmt(0.0) == "0.0");
  CHECK(fmt(-0.0) == "-0.0");
  CHECK(fmt(1.5) == "1.5");
  CHECK(fmt(0.1) == "0.1");
  CHECK(fmt(static_cast<double>(0.1f)) == "0.10000000149011612");
  CHECK(fmt(123456789.0) == "123456789.0");
  CHECK(fmt(1e21) == "1e21");
  CHECK(fmt(1e20) == "100000000000000000000.0");
  CHECK(fmt(1.5e-7) == "1.5e-7");
  CHECK(fmt(0.000001) == "0.000001");
  CHECK(fmt(1e-7) == "1e-7");
  CHECK(fmt(-2.5e30) == "-2.5e30");
}

// response "data" arrays -> tensor bytes, through the public ParseResponseBody
static std::string OutputBytes(const std::string& datatype, const std::string& data, bool expect_ok = true) {
  const std::string js = "{\"model_name\":\"m\",\"model_version\":\"1\",\"outputs\":[{\"name\":\"O\",\"datatype\":\"" + datatype +
                         "\",\"shape\":[2],\"data\":" + data + "}]}";
  tc::InferResult* result = nullptr;
  CHECK_OK(tc::InferenceServerHttpClient::ParseResponseBody(&result, std::vector<char>(js.begin(), js.end())));
  const uint8_t* buf = nullptr;
  size_t n = 0;
  tc::Error err = result->RawData("O", &buf, &n);
  CHECK(err.IsOk() == expect_ok);
  std::string out = err.IsOk() ? std::string(reinterpret_cast<const char*>(buf), n) : std::string();
  delete result;
  return out;
}

template <typename T>
static bool Is(const std::string& bytes, T a, T b) {
  T v[2];
  if (bytes.size() != sizeof(v)) return false;
  memcpy(v, bytes.data(), sizeof(v));
  return v[0] == a && v[1] == b;
}

static void TestJsonOutputToBinary() {
  CHECK(Is<uint8_t>(OutputBytes("BOOL", "[false, true]"), 0, 1));
  CHECK(Is<uint8_t>(OutputBytes("UINT8", "[1, 255]"), 1, 255));
  CHECK(Is<uint16_t>(OutputBytes("UINT16", "[1, 65535]"), 1, 65535));
  CHECK(Is<uint32_t>(OutputBytes("UINT32", "[1, 4294967295]"), 1, 4294967295u));
  CHECK(Is<uint64_t>(OutputBytes("UINT64", "[1, 18446744073709551615]"), 1, 18446744073709551615ULL));
  CHECK(Is<int8_t>(OutputBytes("INT8", "[-128, 127]"), -128, 127));
  CHECK(Is<int16_t>(OutputBytes("INT16", "[-32768, 32767]"), -32768, 32767));
  CHECK(Is<int32_t>(OutputBytes("INT32", "[-2147483648, 2147483647]"), INT32_MIN, INT32_MAX));
  CHECK(Is<int64_t>(OutputBytes("INT64", "[-9223372036854775808, 9223372036854775807]"), INT64_MIN, INT64_MAX));
  OutputBytes("FP16", "[1.0, 2.0]", false);
  CHECK(Is<float>(OutputBytes("FP32", "[-1000.0, 1000.0]"), -1000.0f, 1000.0f));
  CHECK(Is<double>(OutputBytes("FP64", "[-1000.0, 1000.0]"), -1000.0, 1000.0));
  const std::string framed = OutputBytes("BYTES", "[\"ab\", \"cde\"]");
  CHECK(framed == std::string("\x02\0\0\0ab\x03\0\0\0cde", 13));
  OutputBytes("BF16", "[1.0, 2.0]", false);
  OutputBytes("invaliddatatype", "[1, 2]", false);
}

static void TestScatterList() {
  // GetNext in copy mode walks across buffer boundaries (common.cc:245-273)
  tc::InferInput* input = nullptr;
  tc::InferInput::Create(&input, "X", {10}, "UINT8");
  const uint8_t a[4] = {0, 1, 2, 3}, b[6] = {4, 5, 6, 7, 8, 9};
  input->AppendRaw(a, 4);
  input->AppendRaw(b, 6);
  input->PrepareForRequest();
  uint8_t out[10];
  size_t n = 0, total = 0;
  bool end = false;
  input->GetNext(out, 3, &n, &end);
  CHECK(n == 3 && !end);
  total += n;
  input->GetNext(out + total, 3, &n, &end);
  CHECK(n == 3 && !end);
  total += n;
  input->GetNext(out + total, 100, &n, &end);
  CHECK(n == 4 && end);
  for (int i = 0; i < 10; ++i) CHECK(out[i] == i);
  // BYTES framing: <u32 length><payload> (Rust KAT infer.rs:1095-1106: "hello","world" = 18 B)
  tc::InferInput* s = nullptr;
  tc::InferInput::Create(&s, "S", {2}, "BYTES");
  s->AppendFromString({"hello", "world"});
  const uint8_t* buf = nullptr;
  size_t bs = 0;
  s->RawData(&buf, &bs);
  CHECK(bs == 18 && memcmp(buf, "\x05\0\0\0hello\x05\0\0\0world", 18) == 0);
  std::string name;
  size_t sz, off;
  CHECK(!s->SharedMemoryInfo(&name, &sz, &off).IsOk());
  s->SetSharedMemory("region", 64, 8);
  CHECK(s->IsSharedMemory() && s->SharedMemoryInfo(&name, &sz, &off).IsOk() && name == "region" && sz == 64 && off == 8);
  s->Reset();
  CHECK(!s->IsSharedMemory());
  delete input;
  delete s;
}

static void TestRequestBody() {
  // header layout of the reference C++ client: id, parameters, inputs, outputs
  tc::InferInput *in0 = nullptr, *in1 = nullptr;
  tc::InferInput::Create(&in0, "INPUT0", {1, 4}, "INT32");
  tc::InferInput::Create(&in1, "INPUT1", {1, 4}, "INT32");
  int32_t a[4] = {1, 2, 3, 4}, b[4] = {-1, -2, -3, -4};
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a), 8);
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a) + 8, 8);
  in1->AppendRaw(reinterpret_cast<uint8_t*>(b), 16);
  in1->SetBinaryData(false);
  tc::InferRequestedOutput *o0 = nullptr, *o1 = nullptr;
  tc::InferRequestedOutput::Create(&o0, "OUTPUT0");
  tc::InferRequestedOutput::Create(&o1, "OUTPUT1", 3);
  o1->SetSharedMemory("out_region", 16, 32);
  tc::InferOptions options("simple");
  options.request_id_ = "7";
  options.sequence_id_ = 5;
  options.sequence_start_ = true;
  options.priority_ = 2;
  options.request_parameters["k"] = tc::RequestParameter{"k", "v", "string"};
  std::vector<char> body;
  size_t header_length = 0;
  CHECK_OK(tc::InferenceServerHttpClient::GenerateRequestBody(&body, &header_length, options, {in0, in1}, {o0, o1}));
  const std::string want =
      "{\"id\":\"7\",\"parameters\":{\"sequence_id\":5,\"sequence_start\":true,\"sequence_end\":false,\"priority\":2,\"k\":\"v\"},"
      "\"inputs\":[{\"name\":\"INPUT0\",\"datatype\":\"INT32\",\"shape\":[1,4],\"parameters\":{\"binary_data_size\":16}},"
      "{\"name\":\"INPUT1\",\"datatype\":\"INT32\",\"shape\":[1,4],\"data\":[-1,-2,-3,-4]}],"
      "\"outputs\":[{\"name\":\"OUTPUT0\",\"parameters\":{\"binary_data\":true}},"
      "{\"name\":\"OUTPUT1\",\"parameters\":{\"classification\":3,\"shared_memory_region\":\"out_region\","
      "\"shared_memory_byte_size\":16,\"shared_memory_offset\":32}}]}";
  CHECK(std::string(body.data(), header_length) == want);
  CHECK(body.size() == header_length + 16 && memcmp(body.data() + header_length, a, 16) == 0);
  // no outputs requested -> binary_data_output parameter
  tc::InferOptions plain("simple");
  CHECK_OK(tc::InferenceServerHttpClient::GenerateRequestBody(&body, &header_length, plain, {in0}));
  CHECK(std::string(body.data(), header_length) ==
        "{\"id\":\"\",\"parameters\":{\"binary_data_output\":true},\"inputs\":[{\"name\":\"INPUT0\",\"datatype\":\"INT32\","
        "\"shape\":[1,4],\"parameters\":{\"binary_data_size\":16}}]}");
  // response with binary outputs
  const std::string js = "{\"model_name\":\"simple\",\"model_version\":\"1\",\"id\":\"7\",\"outputs\":[{\"name\":\"OUTPUT0\",\"datatype\":\"INT32\","
                         "\"shape\":[1,4],\"parameters\":{\"binary_data_size\":16}}]}";
  std::vector<char> resp(js.begin(), js.end());
  resp.insert(resp.end(), reinterpret_cast<char*>(b), reinterpret_cast<char*>(b) + 16);
  tc::InferResult* result = nullptr;
  CHECK_OK(tc::InferenceServerHttpClient::ParseResponseBody(&result, resp, js.size()));
  std::string name, dt;
  std::vector<int64_t> shape;
  CHECK_OK(result->ModelName(&name));
  CHECK(name == "simple");
  CHECK_OK(result->Id(&name));
  CHECK(name == "7");
  CHECK_OK(result->Shape("OUTPUT0", &shape));
  CHECK((shape == std::vector<int64_t>{1, 4}));
  CHECK_OK(result->Datatype("OUTPUT0", &dt));
  CHECK(dt == "INT32");
  const uint8_t* buf = nullptr;
  size_t n = 0;
  CHECK_OK(result->RawData("OUTPUT0", &buf, &n));
  CHECK(n == 16 && memcmp(buf, b, 16) == 0);
  CHECK(result->RawData("nope", &buf, &n).Message() == "The response does not contain results for output name nope");
  std::vector<std::string> strs;
  CHECK(!result->StringData("OUTPUT0", &strs).IsOk());
  delete result;
  delete in0;
  delete in1;
  delete o0;
  delete o1;
}

static void TestTimers() {
  tc::RequestTimers t;
  using K = tc::RequestTimers::Kind;
  CHECK(t.Duration(K::REQUEST_START, K::REQUEST_END) == UINT64_MAX);
  t.CaptureTimestamp(K::REQUEST_START);
  t.CaptureTimestamp(K::REQUEST_END);
  CHECK(t.Duration(K::REQUEST_START, K::REQUEST_END) < 1000000000ull);
  CHECK(t.Duration(K::REQUEST_END, K::REQUEST_START) == UINT64_MAX || t.Timestamp(K::REQUEST_END) == t.Timestamp(K::REQUEST_START));
}

// ---- loopback cases against a KServe-v2 server with the `simple` model ---------------------

struct Simple {
  std::vector<int32_t> a, b;
  tc::InferInput *in0 = nullptr, *in1 = nullptr;
  Simple(int base) : a(16), b(16) {
    for (int i = 0; i < 16; ++i) {
      a[i] = base + i;
      b[i] = 2 * i + 1;
    }
    tc::InferInput::Create(&in0, "INPUT0", {1, 16}, "INT32");
    tc::InferInput::Create(&in1, "INPUT1", {1, 16}, "INT32");
    in0->AppendRaw(reinterpret_cast<const uint8_t*>(a.data()), 64);
    in1->AppendRaw(reinterpret_cast<const uint8_t*>(b.data()), 64);
  }
  ~Simple() {
    delete in0;
    delete in1;
  }
  void Validate(tc::InferResult* r, bool want_out1 = true) const {
    CHECK_OK(r->RequestStatus());
    const uint8_t* buf = nullptr;
    size_t n = 0;
    CHECK_OK(r->RawData("OUTPUT0", &buf, &n));
    CHECK(n == 64);
    for (int i = 0; n == 64 && i < 16; ++i) CHECK(reinterpret_cast<const int32_t*>(buf)[i] == a[i] + b[i]);
    tc::Error e1 = r->RawData("OUTPUT1", &buf, &n);
    if (want_out1) {
      CHECK_OK(e1);
      for (int i = 0; e1.IsOk() && n == 64 && i < 16; ++i) CHECK(reinterpret_cast<const int32_t*>(buf)[i] == a[i] - b[i]);
    } else {
      CHECK(!e1.IsOk());
    }
  }
};

static void TestLoopback(const std::string& url) {
  std::unique_ptr<tc::InferenceServerHttpClient> client;
  CHECK_OK(tc::InferenceServerHttpClient::Create(&client, url));
  bool live = false, ready = false;
  CHECK_OK(client->IsServerLive(&live));
  CHECK_OK(client->IsServerReady(&ready));
  CHECK(live && ready);
  CHECK_OK(client->IsModelReady(&ready, "simple"));
  CHECK(ready);
  client->IsModelReady(&ready, "no_such_model");
  CHECK(!ready);
  std::string md;
  CHECK_OK(client->ModelMetadata(&md, "simple"));
  CHECK(md.find("\"INPUT0\"") != std::string::npos);
  CHECK(!client->ModelMetadata(&md, "no_such_model").IsOk());

  tc::InferRequestedOutput *o0 = nullptr, *o1 = nullptr;
  tc::InferRequestedOutput::Create(&o0, "OUTPUT0");
  tc::InferRequestedOutput::Create(&o1, "OUTPUT1");
  tc::InferOptions options("simple");
  options.request_id_ = "cc-1";

  // sync, binary in / binary out
  {
    Simple s(10);
    tc::InferResult* r = nullptr;
    CHECK_OK(client->Infer(&r, options, {s.in0, s.in1}, {o0, o1}));
    s.Validate(r);
    std::string id;
    CHECK_OK(r->Id(&id));
    CHECK(id == "cc-1");
    delete r;
  }
  // JSON in (INPUT1) / JSON out (OUTPUT1)
  {
    Simple s(100);
    s.in1->SetBinaryData(false);
    o1->SetBinaryData(false);
    tc::InferResult* r = nullptr;
    CHECK_OK(client->Infer(&r, options, {s.in0, s.in1}, {o0, o1}));
    s.Validate(r);
    delete r;
    o1->SetBinaryData(true);
  }
  // no outputs requested: every output comes back as binary
  {
    Simple s(-5);
    tc::InferResult* r = nullptr;
    CHECK_OK(client->Infer(&r, options, {s.in0, s.in1}));
    s.Validate(r);
    delete r;
  }
  // unknown model: the server's message becomes the Error
  {
    Simple s(0);
    tc::InferOptions bad("no_such_model");
    tc::InferResult* r = nullptr;
    tc::Error err = client->Infer(&r, bad, {s.in0, s.in1});
    CHECK(!err.IsOk() && err.Message().find("no_such_model") != std::string::npos);
    delete r;
  }
  // InferMulti: three requests, different outputs per request (cc_client_test.cc:356-419)
  {
    Simple s0(1), s1(2), s2(3);
    std::vector<tc::InferOptions> opts(3, tc::InferOptions("simple"));
    std::vector<std::vector<tc::InferInput*>> inputs = {{s0.in0, s0.in1}, {s1.in0, s1.in1}, {s2.in0, s2.in1}};
    std::vector<std::vector<const tc::InferRequestedOutput*>> outputs = {{o0, o1}, {o0}, {o0, o1}};
    std::vector<tc::InferResult*> results;
    CHECK_OK(client->InferMulti(&results, opts, inputs, outputs));
    CHECK(results.size() == 3);
    if (results.size() == 3) {
      s0.Validate(results[0]);
      s1.Validate(results[1], false);
      s2.Validate(results[2]);
    }
    for (tc::InferResult* r : results) delete r;
    // one option / one output set shared by all requests
    results.clear();
    CHECK_OK(client->InferMulti(&results, {tc::InferOptions("simple")}, inputs, {{o0, o1}}));
    CHECK(results.size() == 3);
    for (tc::InferResult* r : results) delete r;
    // mismatched counts are rejected before anything is sent (:636-703)
    results.clear();
    std::vector<tc::InferOptions> two(2, tc::InferOptions("simple"));
    CHECK(client->InferMulti(&results, two, inputs, outputs).Message() ==
          "'options' must either contain 1 element or match size of 'inputs'");
    outputs.pop_back();
    CHECK(client->InferMulti(&results, opts, inputs, outputs).Message() ==
          "'outputs' must either contain 0/1 element or match size of 'inputs'");
  }
  // AsyncInfer: callbacks on the worker thread
  {
    Simple s(42);
    std::mutex mu;
    std::condition_variable cv;
    int done = 0;
    for (int i = 0; i < 4; ++i) {
      CHECK_OK(client->AsyncInfer(
          [&](tc::InferResult* r) {
            s.Validate(r);
            delete r;
            std::lock_guard<std::mutex> lk(mu);
            ++done;
            cv.notify_one();
          },
          options, {s.in0, s.in1}, {o0, o1}));
    }
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return done == 4; }));
  }
  // AsyncInferMulti (:705-771)
  {
    Simple s0(7), s1(8);
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    CHECK_OK(client->AsyncInferMulti(
        [&](std::vector<tc::InferResult*> results) {
          CHECK(results.size() == 2);
          if (results.size() == 2) {
            s0.Validate(results[0]);
            s1.Validate(results[1]);
          }
          for (tc::InferResult* r : results) delete r;
          std::lock_guard<std::mutex> lk(mu);
          done = true;
          cv.notify_one();
        },
        {tc::InferOptions("simple")}, {{s0.in0, s0.in1}, {s1.in0, s1.in1}}, {{o0, o1}}));
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return done; }));
  }
  tc::InferStat stat;
  CHECK_OK(client->ClientInferStat(&stat));
  CHECK(stat.completed_request_count >= 12 && stat.cumulative_total_request_time_ns > stat.cumulative_send_time_ns);
  delete o0;
  delete o1;
}

static void TestJsonParser() {
  using tb200::json::Value;
  Value v;
  std::string err;
  const std::string doc = R"({"a":[1,-2,3.5,1e3,18446744073709551615],"s":"x\"y\u00e9\ud83d\ude00\n","o":{"t":true,"f":false,"n":null},"e":[],"eo":{}})";
  CHECK(Value::Parse(doc.data(), doc.size(), &v, &err));
  const Value* a = v.Find("a");
  CHECK(a != nullptr && a->size() == 5);
  int64_t i = 0;
  uint64_t u = 0;
  double d = 0;
  CHECK((*a)[0].AsInt(&i) && i == 1);
  CHECK((*a)[1].AsInt(&i) && i == -2);
  CHECK((*a)[2].AsDouble(&d) && d == 3.5);
  CHECK((*a)[3].AsDouble(&d) && d == 1000.0);
  CHECK((*a)[4].AsUInt(&u) && u == 18446744073709551615ULL);
  CHECK(v.Find("s")->str() == std::string("x\"y\xc3\xa9\xf0\x9f\x98\x80\n"));
  bool b = false;
  CHECK(v.Find("o")->Find("t")->AsBool(&b) && b);
  CHECK(v.Find("o")->Find("n")->is_null());
  CHECK(v.Find("e")->is_array() && v.Find("e")->size() == 0 && v.Find("eo")->is_object());
  // writer round trip keeps member order and escapes
  Value again;
  const std::string dumped = v.Dump();
  CHECK(Value::Parse(dumped.data(), dumped.size(), &again, &err) && again.Dump() == dumped);
  CHECK(dumped.find("\\\"") != std::string::npos && dumped.rfind("{\"a\":[1,-2,3.5,1000.0,18446744073709551615]", 0) == 0);
  for (const char* bad : {"{", "[1,]", "{\"a\" 1}", "tru", "\"abc", "[1] x", "{\"a\":}", ""}) {
    Value w;
    CHECK(!Value::Parse(bad, strlen(bad), &w, &err));
  }
  // an error body and a malformed body through the result class
  tc::InferResult* r = nullptr;
  const std::string junk = "not json";
  tc::InferenceServerHttpClient::ParseResponseBody(&r, std::vector<char>(junk.begin(), junk.end()));
  CHECK(!r->RequestStatus().IsOk());
  std::string name;
  CHECK(!r->ModelName(&name).IsOk());
  delete r;
}

static void TestBytesInputFromStrings() {
  tc::InferInput* s = nullptr;
  tc::InferInput::Create(&s, "S", {1, 2}, "BYTES");
  s->AppendFromString({"ab", "c"});
  std::vector<char> body;
  size_t header_length = 0;
  tc::InferOptions options("m");
  CHECK_OK(tc::InferenceServerHttpClient::GenerateRequestBody(&body, &header_length, options, {s}));
  CHECK(std::string(body.data(), header_length).find("\"binary_data_size\":11") != std::string::npos);
  CHECK(body.size() == header_length + 11 && memcmp(body.data() + header_length, "\x02\0\0\0ab\x01\0\0\0c", 11) == 0);
  s->SetBinaryData(false);
  CHECK_OK(tc::InferenceServerHttpClient::GenerateRequestBody(&body, &header_length, options, {s}));
  CHECK(std::string(body.data(), body.size()).find("\"data\":[\"ab\",\"c\"]") != std::string::npos && body.size() == header_length);
  delete s;
}

// body compression (http_client.cc:146-254): responses inflated with zlib on the host; requests
// deflated by the device encoder (argv[2] says whether a CUDA device is expected)
static void TestCompression(const std::string& url, bool expect_device) {
  using CT = tc::InferenceServerHttpClient::CompressionType;
  std::unique_ptr<tc::InferenceServerHttpClient> client;
  CHECK_OK(tc::InferenceServerHttpClient::Create(&client, url));
  std::vector<int32_t> big(50000);
  for (size_t i = 0; i < big.size(); ++i) big[i] = static_cast<int32_t>(i % 251);
  tc::InferInput* in;
  tc::InferInput::Create(&in, "INPUT0", {static_cast<int64_t>(big.size())}, "INT32");
  in->AppendRaw(reinterpret_cast<uint8_t*>(big.data()), big.size() * 4);
  tc::InferOptions opt("custom_identity_int32");
  auto same = [&](tc::InferResult* r) {
    const uint8_t* p = nullptr;
    size_t n = 0;
    return r != nullptr && r->RequestStatus().IsOk() && r->RawData("OUTPUT0", &p, &n).IsOk() && n == big.size() * 4 &&
           memcmp(p, big.data(), n) == 0;
  };
  for (CT response : {CT::GZIP, CT::DEFLATE}) {  // the server compresses, zlib inflates
    tc::InferResult* result = nullptr;
    CHECK_OK(client->Infer(&result, opt, {in}, {}, tc::Headers(), tc::Parameters(), CT::NONE, response));
    CHECK(same(result));
    delete result;
  }
  {
    std::mutex mu;
    std::condition_variable cv;
    bool done = false, ok = false;
    CHECK_OK(client->AsyncInfer(
        [&](tc::InferResult* r) {
          std::lock_guard<std::mutex> lk(mu);
          ok = same(r);
          done = true;
          delete r;
          cv.notify_all();
        },
        opt, {in}, {}, tc::Headers(), tc::Parameters(), CT::NONE, CT::GZIP));
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return done; }) && ok);
  }
  for (CT request : {CT::GZIP, CT::DEFLATE}) {  // made by tb200_deflate_async, inflated by the server
    tc::InferResult* result = nullptr;
    tc::Error err = client->Infer(&result, opt, {in}, {}, tc::Headers(), tc::Parameters(), request, CT::GZIP);
    if (expect_device) {
      CHECK_OK(err);
      CHECK(same(result));
    } else {
      CHECK(!err.IsOk() && err.Message().find("no CUDA device") != std::string::npos);
    }
    delete result;
  }
  std::string plain;
  CHECK(!tc::detail::Inflate("not a zlib stream", &plain).IsOk());
  delete in;
}

int main(int argc, char** argv) {
  TestJsonParser();
  TestBytesInputFromStrings();
  TestBinaryInputsToJson();
  TestBinaryInputToJson();
  TestDoubleFormatting();
  TestJsonOutputToBinary();
  TestScatterList();
  TestRequestBody();
  TestTimers();
  if (argc > 2 && std::string(argv[2]).rfind("compress", 0) == 0) TestCompression(argv[1], std::string(argv[2]) == "compress-gpu");
  else if (argc > 1) TestLoopback(argv[1]);
  if (g_failures == 0) {
    std::cout << "PASS" << (argc > 1 ? " (offline + loopback)" : " (offline)") << std::endl;
    return 0;
  }
  std::cout << "FAIL: " << g_failures << " check(s)" << std::endl;
  return 1;
}
