// C++ gRPC front end (client_b200/cpp/tb200_grpc_client.h): message codec known answers and
// round trips, request serialisation (checked by the caller against libprotobuf via Python),
// and -- with a server URL -- the reference's gRPC client behaviours against the mock server:
// health / metadata, Infer with AppendRaw scatter lists, errors as trailers-only responses,
// AsyncInfer / InferMulti / AsyncInferMulti (cc_client_test.cc:169-1032 shapes), BYTES, the
// bidirectional stream, client time-outs.
//
//   test_cc_grpc_client                      offline checks
//   test_cc_grpc_client --roundtrip FILE     lines "<Type> <hex>" -> "<hex of re-serialisation>\n<DebugString>\n---"
//   test_cc_grpc_client --requests           hex of SerializeInferRequest() for fixed calls
//   test_cc_grpc_client HOST:PORT [slow]     loopback checks ("slow": the server sleeps 600 ms per request)
//   test_cc_grpc_client HOST:PORT compress-gpu|compress-nogpu   request compression on the device / its refusal without one
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>

#include "grpc_client.h"

namespace tc = triton::client;

static int g_failures = 0;
#define CHECK(cond)                                                                   \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ++g_failures;                                                                   \
      std::cout << "CHECK failed at line " << __LINE__ << ": " #cond << std::endl;    \
    }                                                                                 \
  } while (0)
#define CHECK_OK(expr)                                                                               \
  do {                                                                                               \
    tc::Error e__ = (expr);                                                                          \
    if (!e__.IsOk()) {                                                                               \
      ++g_failures;                                                                                  \
      std::cout << "error at line " << __LINE__ << ": " << e__.Message() << std::endl;               \
    }                                                                                                \
  } while (0)

static std::string Hex(const std::string& s) {
  static const char* d = "0123456789abcdef";
  std::string out;
  for (unsigned char c : s) {
    out.push_back(d[c >> 4]);
    out.push_back(d[c & 15]);
  }
  return out;
}
static std::string Unhex(const std::string& h) {
  std::string out;
  for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back(static_cast<char>(std::stoi(h.substr(i, 2), nullptr, 16)));
  return out;
}

static void TestCodec() {
  // {model_name:"m", raw_input_contents:[01 02]} = 0a016d3a020102
  inference::ModelInferRequest r;
  r.set_model_name("m");
  r.add_raw_input_contents("\x01\x02", 2);
  CHECK(Hex(r.SerializeAsString()) == "0a016d3a020102");
  // proto3 defaults are not serialised; oneof members are, even when zero
  inference::InferParameter p;
  CHECK(p.SerializeAsString().empty() && p.parameter_choice_case() == inference::InferParameter::PARAMETER_CHOICE_NOT_SET);
  p.set_bool_param(false);
  CHECK(Hex(p.SerializeAsString()) == "0800" && p.has_bool_param() && p.parameter_choice_case() == inference::InferParameter::kBoolParam);
  p.set_int64_param(-1);
  CHECK(Hex(p.SerializeAsString()) == "10ffffffffffffffffff01" && !p.has_bool_param() && p.int64_param() == -1 && p.bool_param() == false);
  // packed repeated scalars, negative int32 as 10-byte varint, floats little-endian
  inference::InferTensorContents c;
  c.add_int_contents(1);
  c.add_int_contents(-2);
  c.add_fp32_contents(1.5f);
  c.add_bool_contents(true);
  CHECK(Hex(c.SerializeAsString()) == "0a0101120b01feffffffffffffffff0132040000c03f");
  inference::InferTensorContents d;
  CHECK(d.ParseFromString(c.SerializeAsString()) && d.int_contents_size() == 2 && d.int_contents(1) == -2 && d.fp32_contents(0) == 1.5f &&
        d.bool_contents(0));
  // unpacked encoding of a packed field is accepted, unknown fields are skipped
  CHECK(d.ParseFromString(Unhex("10031005" "7a03616263")) && d.int_contents_size() == 2 && d.int_contents(1) == 5);
  // truncated input is an error
  CHECK(!d.ParseFromString(Unhex("0a05")));
  // maps, nested messages, text format
  inference::ModelInferRequest q;
  q.set_model_name("simple");
  auto* in = q.add_inputs();
  in->set_name("INPUT0");
  in->set_datatype("INT32");
  in->add_shape(1);
  in->add_shape(16);
  (*in->mutable_parameters())["shared_memory_byte_size"].set_int64_param(64);
  (*q.mutable_parameters())["sequence_start"].set_bool_param(true);
  inference::ModelInferRequest q2;
  CHECK(q2.ParseFromString(q.SerializeAsString()) && q2.SerializeAsString() == q.SerializeAsString());
  CHECK(q2.inputs(0).parameters().at("shared_memory_byte_size").int64_param() == 64);
  CHECK(q2.DebugString() ==
        "model_name: \"simple\"\nparameters {\n  key: \"sequence_start\"\n  value {\n    bool_param: true\n  }\n}\ninputs {\n  name: \"INPUT0\"\n"
        "  datatype: \"INT32\"\n  shape: 1\n  shape: 16\n  parameters {\n    key: \"shared_memory_byte_size\"\n    value {\n      int64_param: 64\n"
        "    }\n  }\n}\n");
  // copies are deep
  inference::ModelConfigResponse cfg;
  cfg.mutable_config()->set_name("a");
  inference::ModelConfigResponse cfg2 = cfg;
  cfg2.mutable_config()->set_name("b");
  CHECK(cfg.config().name() == "a" && cfg2.config().name() == "b");
  inference::ModelConfigResponse empty;
  CHECK(!empty.has_config() && empty.config().name().empty());
}

static void TestResultDecoding() {
  // InferResult over a response: grpc_client.cc:178-446 behaviours
  std::string message;
  tc::InferInput* in0;
  tc::InferInput::Create(&in0, "INPUT0", {1, 4}, "INT32");
  int32_t a[2] = {1, 2}, b[2] = {3, 4};
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a), sizeof(a));
  in0->AppendRaw(reinterpret_cast<uint8_t*>(b), sizeof(b));  // scatter list of two buffers -> one raw_input_contents entry
  tc::InferOptions options("simple");
  CHECK_OK(tc::InferenceServerGrpcClient::SerializeInferRequest(&message, options, {in0}));
  inference::ModelInferRequest parsed;
  CHECK(parsed.ParseFromString(message) && parsed.raw_input_contents_size() == 1 && parsed.raw_input_contents(0).size() == 16);
  int32_t got[4];
  memcpy(got, parsed.raw_input_contents(0).data(), 16);
  CHECK(got[0] == 1 && got[1] == 2 && got[2] == 3 && got[3] == 4);
  CHECK(parsed.parameters().count("triton_enable_empty_final_response") == 1);
  // shared-memory inputs carry parameters and no raw contents
  in0->SetSharedMemory("region", 16, 32);
  CHECK_OK(tc::InferenceServerGrpcClient::SerializeInferRequest(&message, options, {in0}));
  CHECK(parsed.ParseFromString(message) && parsed.raw_input_contents_size() == 0);
  CHECK(parsed.inputs(0).parameters().at("shared_memory_region").string_param() == "region" &&
        parsed.inputs(0).parameters().at("shared_memory_offset").int64_param() == 32);
  delete in0;
}

static void PrintRequests() {
  // fixed calls; the caller rebuilds the same messages with libprotobuf (Python) and compares bytes
  int32_t x[16], y[16];
  for (int i = 0; i < 16; ++i) {
    x[i] = i;
    y[i] = -1;
  }
  tc::InferInput *in0, *in1;
  tc::InferInput::Create(&in0, "INPUT0", {1, 16}, "INT32");
  tc::InferInput::Create(&in1, "INPUT1", {1, 16}, "INT32");
  in0->AppendRaw(reinterpret_cast<uint8_t*>(x), sizeof(x));
  in1->AppendRaw(reinterpret_cast<uint8_t*>(y), 32);
  in1->AppendRaw(reinterpret_cast<uint8_t*>(y) + 32, 32);
  tc::InferRequestedOutput *o0, *o1;
  tc::InferRequestedOutput::Create(&o0, "OUTPUT0");
  tc::InferRequestedOutput::Create(&o1, "OUTPUT1", 3);
  std::string m;
  {
    tc::InferOptions opt("simple");
    tc::InferenceServerGrpcClient::SerializeInferRequest(&m, opt, {in0, in1}, {o0, o1});
    std::cout << "plain " << Hex(m) << std::endl;
  }
  {
    tc::InferOptions opt("simple");
    opt.model_version_ = "2";
    opt.request_id_ = "req-7";
    opt.sequence_id_ = 1007;
    opt.sequence_start_ = true;
    opt.priority_ = 3;
    opt.server_timeout_ = 5000;
    opt.triton_enable_empty_final_response_ = true;
    opt.request_parameters["my_key"] = tc::RequestParameter{"my_key", "v", "string"};
    opt.request_parameters["count"] = tc::RequestParameter{"count", "-12", "int"};
    opt.request_parameters["flag"] = tc::RequestParameter{"flag", "true", "bool"};
    tc::InferenceServerGrpcClient::SerializeInferRequest(&m, opt, {in0, in1}, {o0});
    std::cout << "options " << Hex(m) << std::endl;
  }
  {
    tc::InferOptions opt("simple");
    opt.sequence_id_str_ = "seq-a";
    opt.sequence_end_ = true;
    in0->SetSharedMemory("input_data", 64);
    in1->SetSharedMemory("input_data", 64, 64);
    o0->SetSharedMemory("output_data", 64);
    o1->SetSharedMemory("output_data", 64, 64);
    tc::InferenceServerGrpcClient::SerializeInferRequest(&m, opt, {in0, in1}, {o0, o1});
    std::cout << "shm " << Hex(m) << std::endl;
  }
  delete in0;
  delete in1;
  delete o0;
  delete o1;
}

template <typename T>
static bool RoundTripAs(const std::string& bytes) {
  T m;
  const bool ok = m.ParseFromString(bytes);
  std::cout << (ok ? Hex(m.SerializeAsString()) : std::string("PARSE-ERROR")) << "\n" << m.DebugString() << "---" << std::endl;
  return ok;
}
static int RoundTrip(const char* path) {
  std::ifstream f(path);
  std::string type, hex;
  while (f >> type >> hex) {
    if (hex == "-") hex.clear();
    const std::string bytes = Unhex(hex);
    if (type == "ModelInferRequest") RoundTripAs<inference::ModelInferRequest>(bytes);
    else if (type == "ModelInferResponse") RoundTripAs<inference::ModelInferResponse>(bytes);
    else if (type == "ModelStreamInferResponse") RoundTripAs<inference::ModelStreamInferResponse>(bytes);
    else if (type == "ModelConfigResponse") RoundTripAs<inference::ModelConfigResponse>(bytes);
    else if (type == "ModelMetadataResponse") RoundTripAs<inference::ModelMetadataResponse>(bytes);
    else if (type == "ModelStatisticsResponse") RoundTripAs<inference::ModelStatisticsResponse>(bytes);
    else if (type == "CudaSharedMemoryStatusResponse") RoundTripAs<inference::CudaSharedMemoryStatusResponse>(bytes);
    else if (type == "InferTensorContents") RoundTripAs<inference::InferTensorContents>(bytes);
    else if (type == "RepositoryModelLoadRequest") RoundTripAs<inference::RepositoryModelLoadRequest>(bytes);
    else if (type == "LogSettingsRequest") RoundTripAs<inference::LogSettingsRequest>(bytes);
    else {
      std::cout << "UNKNOWN-TYPE\n---" << std::endl;
    }
  }
  return 0;
}

// ---- loopback -------------------------------------------------------------------------------
static void AddSubInputs(std::vector<int32_t>* a, std::vector<int32_t>* b, tc::InferInput** in0, tc::InferInput** in1) {
  a->resize(16);
  b->resize(16);
  for (int i = 0; i < 16; ++i) {
    (*a)[i] = i;
    (*b)[i] = 1;
  }
  tc::InferInput::Create(in0, "INPUT0", {1, 16}, "INT32");
  tc::InferInput::Create(in1, "INPUT1", {1, 16}, "INT32");
  // two borrowed buffers per input (common.h:274-293)
  (*in0)->AppendRaw(reinterpret_cast<uint8_t*>(a->data()), 24);
  (*in0)->AppendRaw(reinterpret_cast<uint8_t*>(a->data()) + 24, 40);
  (*in1)->AppendRaw(reinterpret_cast<uint8_t*>(b->data()), 64);
}
static bool AddSubOk(tc::InferResult* r, const std::vector<int32_t>& a, const std::vector<int32_t>& b) {
  if (r == nullptr || !r->RequestStatus().IsOk()) return false;
  const uint8_t *p0, *p1;
  size_t n0, n1;
  if (!r->RawData("OUTPUT0", &p0, &n0).IsOk() || !r->RawData("OUTPUT1", &p1, &n1).IsOk() || n0 != 64 || n1 != 64) return false;
  for (int i = 0; i < 16; ++i) {
    int32_t s, d;
    memcpy(&s, p0 + 4 * i, 4);
    memcpy(&d, p1 + 4 * i, 4);
    if (s != a[i] + b[i] || d != a[i] - b[i]) return false;
  }
  return true;
}

static void TestLoopback(const std::string& url) {
  std::unique_ptr<tc::InferenceServerGrpcClient> client;
  CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, url));
  bool live = false, ready = false, model_ready = false;
  CHECK_OK(client->IsServerLive(&live));
  CHECK_OK(client->IsServerReady(&ready));
  CHECK_OK(client->IsModelReady(&model_ready, "simple"));
  CHECK(live && ready && model_ready);
  inference::ServerMetadataResponse server;
  CHECK_OK(client->ServerMetadata(&server));
  CHECK(server.name() == "triton" && server.extensions_size() > 0);
  inference::ModelMetadataResponse meta;
  CHECK_OK(client->ModelMetadata(&meta, "simple"));
  CHECK(meta.name() == "simple" && meta.inputs_size() == 2 && meta.inputs(0).datatype() == "INT32" && meta.inputs(0).shape_size() == 2);
  inference::ModelConfigResponse config;
  CHECK_OK(client->ModelConfig(&config, "simple"));
  CHECK(config.config().name() == "simple" && config.config().input_size() == 2 && config.config().input(0).data_type() == inference::TYPE_INT32);
  inference::RepositoryIndexResponse index;
  CHECK_OK(client->ModelRepositoryIndex(&index));
  CHECK(index.models_size() > 3);
  // errors come back as trailers-only responses: the grpc-message is the Error text
  tc::Error err = client->ModelMetadata(&meta, "no_such_model");
  CHECK(!err.IsOk() && err.Message().find("no_such_model") != std::string::npos);

  std::vector<int32_t> a, b;
  tc::InferInput *in0, *in1;
  AddSubInputs(&a, &b, &in0, &in1);
  tc::InferRequestedOutput *o0, *o1;
  tc::InferRequestedOutput::Create(&o0, "OUTPUT0");
  tc::InferRequestedOutput::Create(&o1, "OUTPUT1");
  tc::InferOptions options("simple");
  options.request_id_ = "abc";
  tc::InferResult* result = nullptr;
  CHECK_OK(client->Infer(&result, options, {in0, in1}, {o0, o1}, {{"X-Custom", "1"}}));
  CHECK(AddSubOk(result, a, b));
  std::string name, id, datatype;
  std::vector<int64_t> shape;
  CHECK_OK(result->ModelName(&name));
  CHECK_OK(result->Id(&id));
  CHECK_OK(result->Shape("OUTPUT0", &shape));
  CHECK_OK(result->Datatype("OUTPUT1", &datatype));
  CHECK(name == "simple" && id == "abc" && shape == std::vector<int64_t>({1, 16}) && datatype == "INT32");
  CHECK(!result->Shape("nope", &shape).IsOk() && !result->RawData("nope", nullptr, nullptr).IsOk());
  std::vector<std::string> strings;
  CHECK(result->StringData("OUTPUT0", &strings).Message().find("datatype 'BYTES'") != std::string::npos);
  CHECK(result->DebugString().find("model_name: \"simple\"") != std::string::npos);
  delete result;

  // inference on an unknown model: Infer returns the error AND a result that carries it
  tc::InferOptions bad("no_such_model");
  result = nullptr;
  err = client->Infer(&result, bad, {in0, in1});
  CHECK(!err.IsOk() && result != nullptr && !result->RequestStatus().IsOk());
  delete result;

  // AsyncInfer: callbacks on the client's worker thread, results owned by the callee
  {
    std::mutex mu;
    std::condition_variable cv;
    int done = 0, good = 0;
    for (int i = 0; i < 8; ++i) {
      CHECK_OK(client->AsyncInfer(
          [&](tc::InferResult* r) {
            std::lock_guard<std::mutex> lk(mu);
            ++done;
            if (AddSubOk(r, a, b)) ++good;
            delete r;
            cv.notify_all();
          },
          options, {in0, in1}, {o0, o1}));
    }
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return done == 8; }));
    CHECK(good == 8);
  }
  CHECK(!client->AsyncInfer(nullptr, options, {in0, in1}).IsOk());

  // InferMulti / AsyncInferMulti (cc_client_test.cc:169-1032): one option for all, outputs 0/1/N
  {
    std::vector<tc::InferResult*> results;
    CHECK_OK(client->InferMulti(&results, {options}, {{in0, in1}, {in0, in1}, {in0, in1}}, {{o0, o1}}));
    CHECK(results.size() == 3);
    for (tc::InferResult* r : results) {
      CHECK(AddSubOk(r, a, b));
      delete r;
    }
    results.clear();
    CHECK(client->InferMulti(&results, {options, options}, {{in0, in1}, {in0, in1}, {in0, in1}}).Message() ==
          "'options' must either contain 1 element or match size of 'inputs'");
    CHECK(client->InferMulti(&results, {options}, {{in0, in1}, {in0, in1}, {in0, in1}}, {{o0}, {o1}}).Message() ==
          "'outputs' must either contain 0/1 element or match size of 'inputs'");
    std::mutex mu;
    std::condition_variable cv;
    bool called = false;
    size_t count = 0, good = 0;
    CHECK_OK(client->AsyncInferMulti(
        [&](std::vector<tc::InferResult*> rs) {
          std::lock_guard<std::mutex> lk(mu);
          count = rs.size();
          for (tc::InferResult* r : rs) {
            if (AddSubOk(r, a, b)) ++good;
            delete r;
          }
          called = true;
          cv.notify_all();
        },
        {options}, {{in0, in1}, {in0, in1}}));
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return called; }));
    CHECK(count == 2 && good == 2);
  }

  // BYTES
  {
    tc::InferInput* sin;
    tc::InferInput::Create(&sin, "INPUT0", {1, 8}, "BYTES");
    std::vector<std::string> words = {"a", "bc", "", "defg", "h", "ij", "klm", "nopqrstu"};
    CHECK_OK(sin->AppendFromString(words));
    tc::InferOptions sopt("string_identity");
    result = nullptr;
    CHECK_OK(client->Infer(&result, sopt, {sin}));
    std::vector<std::string> back;
    CHECK_OK(result->StringData("OUTPUT0", &back));
    CHECK(back == words);
    delete result;
    delete sin;
  }

  // a message larger than the default HTTP/2 windows (65,535 B) in both directions
  {
    std::vector<int32_t> big(200000);
    for (size_t i = 0; i < big.size(); ++i) big[i] = static_cast<int32_t>(i * 7);
    tc::InferInput* bin;
    tc::InferInput::Create(&bin, "INPUT0", {static_cast<int64_t>(big.size())}, "INT32");
    bin->AppendRaw(reinterpret_cast<uint8_t*>(big.data()), big.size() * 4);
    tc::InferOptions bopt("custom_identity_int32");
    result = nullptr;
    CHECK_OK(client->Infer(&result, bopt, {bin}));
    const uint8_t* p;
    size_t n = 0;
    CHECK_OK(result->RawData("OUTPUT0", &p, &n));
    CHECK(n == big.size() * 4 && memcmp(p, big.data(), n) == 0);
    delete result;
    delete bin;
  }

  // statistics (common.cc:56-106): 1 + 1 failed? no -- only completed requests with all timestamps count
  tc::InferStat stat;
  CHECK_OK(client->ClientInferStat(&stat));
  CHECK(stat.completed_request_count >= 1 + 8 + 3 + 2 + 2 && stat.cumulative_total_request_time_ns > 0);

  // bidirectional stream: sequence of three requests, then a decoupled model that answers N times
  {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int32_t> values;
    int finals = 0, errors = 0;
    CHECK_OK(client->StartStream([&](tc::InferResult* r) {
      std::lock_guard<std::mutex> lk(mu);
      if (!r->RequestStatus().IsOk()) {
        ++errors;
      } else {
        const uint8_t* p;
        size_t n;
        if (r->RawData("OUTPUT", &p, &n).IsOk() && n == 4) {
          int32_t v;
          memcpy(&v, p, 4);
          values.push_back(v);
        }
        bool fin = false;
        r->IsFinalResponse(&fin);
        if (fin) ++finals;
      }
      delete r;
      cv.notify_all();
    }));
    CHECK(client->StartStream([](tc::InferResult*) {}).Message().find("cannot start another stream") == 0);
    tc::InferInput* sin;
    tc::InferInput::Create(&sin, "INPUT", {1, 1}, "INT32");
    int32_t inputs[3] = {0, 5, 7};
    for (int i = 0; i < 3; ++i) {
      sin->Reset();
      sin->AppendRaw(reinterpret_cast<uint8_t*>(&inputs[i]), 4);
      tc::InferOptions sopt("simple_sequence");
      sopt.sequence_id_ = 42;
      sopt.sequence_start_ = i == 0;
      sopt.sequence_end_ = i == 2;
      CHECK_OK(client->AsyncStreamInfer(sopt, {sin}));
    }
    {
      std::unique_lock<std::mutex> lk(mu);
      CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return values.size() == 3 || errors > 0; }));
    }
    CHECK_OK(client->StopStream());
    CHECK(errors == 0 && values.size() == 3);
    if (values.size() == 3) CHECK(values[0] == 1 && values[1] == 5 && values[2] == 7);  // mock: input (+1 on sequence start)
    CHECK(client->AsyncStreamInfer(tc::InferOptions("simple_sequence"), {sin}).Message() == "Stream has been closed.");
    // decoupled model: one request, N responses, then the empty final response that
    // triton_enable_empty_final_response_ asks for (IsNullResponse)
    {
      std::vector<int32_t> outs;
      int nulls = 0;
      CHECK_OK(client->StartStream([&](tc::InferResult* r) {
        std::lock_guard<std::mutex> lk(mu);
        bool is_null = false, is_final = false;
        r->IsNullResponse(&is_null);
        r->IsFinalResponse(&is_final);
        const uint8_t* p;
        size_t n;
        if (is_null && is_final) ++nulls;
        else if (r->RequestStatus().IsOk() && r->RawData("OUT", &p, &n).IsOk() && n == 4 && !is_final) {
          int32_t v;
          memcpy(&v, p, 4);
          outs.push_back(v);
        }
        delete r;
        cv.notify_all();
      }, false));
      tc::InferInput* rin;
      tc::InferInput::Create(&rin, "IN", {4}, "INT32");
      int32_t reps[4] = {4, 5, 6, 7};
      rin->AppendRaw(reinterpret_cast<uint8_t*>(reps), sizeof(reps));
      tc::InferOptions ropt("repeat_int32");
      ropt.triton_enable_empty_final_response_ = true;
      CHECK_OK(client->AsyncStreamInfer(ropt, {rin}));
      {
        std::unique_lock<std::mutex> lk(mu);
        CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return nulls == 1; }));
      }
      CHECK_OK(client->StopStream());
      CHECK(outs == std::vector<int32_t>({4, 5, 6, 7}) && nulls == 1);
      delete rin;
    }
    // an error inside the stream arrives through the callback, the stream stays usable
    errors = 0;
    CHECK_OK(client->StartStream([&](tc::InferResult* r) {
      std::lock_guard<std::mutex> lk(mu);
      if (!r->RequestStatus().IsOk()) ++errors;
      delete r;
      cv.notify_all();
    }));
    CHECK_OK(client->AsyncStreamInfer(tc::InferOptions("no_such_model"), {sin}));
    {
      std::unique_lock<std::mutex> lk(mu);
      CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return errors == 1; }));
    }
    CHECK_OK(client->StopStream());
    delete sin;
  }

  delete in0;
  delete in1;
  delete o0;
  delete o1;
}

static void TestSlowServer(const std::string& url) {
  // client_timeout_ (microseconds) -> grpc-timeout + local deadline; the reference reports
  // "Deadline Exceeded" (client_timeout_test.cc)
  std::unique_ptr<tc::InferenceServerGrpcClient> client;
  CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, url, false, false, tc::SslOptions(), tc::KeepAliveOptions(), false));
  std::vector<int32_t> a, b;
  tc::InferInput *in0, *in1;
  AddSubInputs(&a, &b, &in0, &in1);
  tc::InferOptions options("simple");
  options.client_timeout_ = 50000;
  tc::InferResult* result = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  tc::Error err = client->Infer(&result, options, {in0, in1});
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  CHECK(!err.IsOk() && err.Message() == "Deadline Exceeded" && ms < 450.0);
  delete result;
  options.client_timeout_ = 5000000;
  result = nullptr;
  CHECK_OK(client->Infer(&result, options, {in0, in1}));
  CHECK(AddSubOk(result, a, b));
  delete result;
  // a client destroyed with asynchronous calls in flight: they are cancelled, their callbacks
  // run (with an error) before the destructor returns, nothing touches the object afterwards
  {
    std::unique_ptr<tc::InferenceServerGrpcClient> doomed, survivor;
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&doomed, url));      // cached channel, shared ...
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&survivor, url));    // ... with this one
    std::atomic<int> called{0}, failed{0};
    tc::InferOptions slow("simple");
    for (int i = 0; i < 3; ++i) {
      CHECK_OK(doomed->AsyncInfer(
          [&](tc::InferResult* r) {
            if (!r->RequestStatus().IsOk()) ++failed;
            ++called;
            delete r;
          },
          slow, {in0, in1}));
    }
    const auto t1 = std::chrono::steady_clock::now();
    doomed.reset();
    const double gone_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    CHECK(called == 3 && failed == 3 && gone_ms < 450.0);
    result = nullptr;
    CHECK_OK(survivor->Infer(&result, slow, {in0, in1}));  // the shared connection is still good
    CHECK(AddSubOk(result, a, b));
    delete result;
  }
  delete in0;
  delete in1;
}

// request compression: the message body becomes a zlib / gzip stream made by the device encoder
// (tb200_deflate_async), flagged in the 5-byte prefix and announced as grpc-encoding; the grpcio
// server inflates it.  Without a device the call reports an Error (there is no host encoder).
static void TestCompression(const std::string& url, bool expect_device, bool responses_only = false) {
  std::unique_ptr<tc::InferenceServerGrpcClient> client;
  CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, url));
  std::vector<int32_t> big(100000);
  for (size_t i = 0; i < big.size(); ++i) big[i] = static_cast<int32_t>(i % 97);
  tc::InferInput* in;
  tc::InferInput::Create(&in, "INPUT0", {static_cast<int64_t>(big.size())}, "INT32");
  in->AppendRaw(reinterpret_cast<uint8_t*>(big.data()), big.size() * 4);
  tc::InferOptions opt("custom_identity_int32");
  if (responses_only) {  // a server that compresses its responses (flag 1 + grpc-encoding): zlib inflates them
    for (int i = 0; i < 3; ++i) {
      tc::InferResult* result = nullptr;
      CHECK_OK(client->Infer(&result, opt, {in}));
      const uint8_t* p = nullptr;
      size_t n = 0;
      if (result != nullptr) {
        CHECK_OK(result->RawData("OUTPUT0", &p, &n));
        CHECK(n == big.size() * 4 && memcmp(p, big.data(), n) == 0);
      }
      delete result;
    }
    delete in;
    return;
  }
  for (grpc_compression_algorithm algo : {GRPC_COMPRESS_GZIP, GRPC_COMPRESS_DEFLATE}) {
    tc::InferResult* result = nullptr;
    tc::Error err = client->Infer(&result, opt, {in}, {}, tc::Headers(), algo);
    if (expect_device) {
      CHECK_OK(err);
      const uint8_t* p = nullptr;
      size_t n = 0;
      if (result != nullptr) {
        CHECK_OK(result->RawData("OUTPUT0", &p, &n));
        CHECK(n == big.size() * 4 && memcmp(p, big.data(), n) == 0);
      }
    } else {
      CHECK(!err.IsOk() && err.Message().find("no CUDA device") != std::string::npos);
    }
    delete result;
  }
  delete in;
}

// The canned-response gRPC stub of libtb200 in this process: a server that goes away fails the
// calls in flight with an Error and the next call opens a new connection (grpc_client.cc keeps
// its channel; here a broken one is replaced, cached or not).
extern "C" {
struct tb200_grpc_stub_server;
int tb200_grpc_stub_server_start(const char* host, int* port, const uint8_t* response, uint64_t response_bytes,
                                 tb200_grpc_stub_server** out);
int tb200_grpc_stub_server_stop(tb200_grpc_stub_server* s);
}
static void TestReconnect() {
  inference::ModelInferResponse canned;
  canned.set_model_name("stub");
  auto* o = canned.add_outputs();
  o->set_name("OUTPUT0");
  o->set_datatype("INT32");
  o->add_shape(2);
  const int32_t vals[2] = {41, 42};
  canned.add_raw_output_contents(vals, sizeof(vals));
  const std::string bytes = canned.SerializeAsString();
  int port = 0;
  tb200_grpc_stub_server* srv = nullptr;
  CHECK(tb200_grpc_stub_server_start("127.0.0.1", &port, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), &srv) == 0);
  const std::string url = "127.0.0.1:" + std::to_string(port);
  std::unique_ptr<tc::InferenceServerGrpcClient> client, second;
  CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, url));
  CHECK_OK(tc::InferenceServerGrpcClient::Create(&second, url));
  tc::InferInput* in;
  tc::InferInput::Create(&in, "INPUT0", {2}, "INT32");
  in->AppendRaw(reinterpret_cast<const uint8_t*>(vals), sizeof(vals));
  tc::InferOptions opt("anything");
  auto infer_ok = [&](tc::InferenceServerGrpcClient* c) {
    tc::InferResult* r = nullptr;
    tc::Error e = c->Infer(&r, opt, {in});
    bool ok = e.IsOk();
    const uint8_t* p = nullptr;
    size_t n = 0;
    if (ok) ok = r->RawData("OUTPUT0", &p, &n).IsOk() && n == 8 && memcmp(p, vals, 8) == 0;
    delete r;
    return ok;
  };
  CHECK(infer_ok(client.get()) && infer_ok(second.get()));
  CHECK(client->GetNumCachedChannels() >= 1);  // both clients share the URL's channel
  tb200_grpc_stub_server_stop(srv);
  CHECK(!infer_ok(client.get()));              // the server is gone: an Error, not a hang
  int again = port;
  srv = nullptr;
  // (the port was an ephemeral one: another socket of this machine may have taken it meanwhile)
  const bool rebound = tb200_grpc_stub_server_start("127.0.0.1", &again, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), &srv) == 0;
  if (!rebound) {
    std::cout << "note: port " << port << " was taken before the stub could return to it; reconnect check skipped" << std::endl;
    delete in;
    return;
  }
  CHECK(infer_ok(client.get()) && infer_ok(second.get()));  // new connection, same client objects
  // many requests in flight on one connection (stream ids advance by two each)
  std::mutex mu;
  std::condition_variable cv;
  int done = 0, good = 0;
  for (int i = 0; i < 200; ++i) {
    CHECK_OK(client->AsyncInfer(
        [&](tc::InferResult* r) {
          std::lock_guard<std::mutex> lk(mu);
          const uint8_t* p = nullptr;
          size_t n = 0;
          if (r->RequestStatus().IsOk() && r->RawData("OUTPUT0", &p, &n).IsOk() && n == 8) ++good;
          ++done;
          delete r;
          cv.notify_all();
        },
        opt, {in}));
  }
  {
    std::unique_lock<std::mutex> lk(mu);
    CHECK(cv.wait_for(lk, std::chrono::seconds(20), [&] { return done == 200; }));
  }
  CHECK(good == 200);
  client.reset();
  second.reset();
  tb200_grpc_stub_server_stop(srv);
  delete in;
}

// KeepAliveOptions (grpc_client.h:63-85): PINGs keep an idle connection checked; a peer that stops
// answering fails the calls in flight with the watchdog's message instead of hanging them.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>
static void TestKeepAlive() {
  inference::ModelInferResponse canned;
  canned.set_model_name("stub");
  const std::string bytes = canned.SerializeAsString();
  int port = 0;
  tb200_grpc_stub_server* srv = nullptr;
  CHECK(tb200_grpc_stub_server_start("127.0.0.1", &port, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size(), &srv) == 0);
  tc::KeepAliveOptions ka;
  ka.keepalive_time_ms = 20;
  ka.keepalive_timeout_ms = 500;
  ka.keepalive_permit_without_calls = true;
  {
    std::unique_ptr<tc::InferenceServerGrpcClient> client;
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, "127.0.0.1:" + std::to_string(port), false, false, tc::SslOptions(), ka, false));
    tc::InferInput* in;
    tc::InferInput::Create(&in, "INPUT0", {1}, "INT32");
    const int32_t one = 1;
    in->AppendRaw(reinterpret_cast<const uint8_t*>(&one), 4);
    for (int i = 0; i < 6; ++i) {  // 60 ms of silence between calls: pings go out and are acknowledged
      tc::InferResult* r = nullptr;
      CHECK_OK(client->Infer(&r, tc::InferOptions("m"), {in}));
      delete r;
      std::this_thread::sleep_for(std::chrono::milliseconds(60));
    }
    delete in;
  }
  tb200_grpc_stub_server_stop(srv);
  // a peer that accepts the connection and never says anything
  const int lfd = socket(AF_INET, SOCK_STREAM, 0);
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  inet_pton(AF_INET, "127.0.0.1", &addr.sin_addr);
  CHECK(bind(lfd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) == 0 && listen(lfd, 4) == 0);
  socklen_t len = sizeof(addr);
  getsockname(lfd, reinterpret_cast<sockaddr*>(&addr), &len);
  std::atomic<int> accepted{-1};
  std::thread mute([&] { accepted = accept(lfd, nullptr, nullptr); });
  ka.keepalive_time_ms = 50;
  ka.keepalive_timeout_ms = 150;
  {
    std::unique_ptr<tc::InferenceServerGrpcClient> client;
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, "127.0.0.1:" + std::to_string(ntohs(addr.sin_port)), false, false, tc::SslOptions(), ka, false));
    bool live = true;
    const auto t0 = std::chrono::steady_clock::now();
    tc::Error err = client->IsServerLive(&live);  // no deadline: only the watchdog can end it
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    CHECK(!err.IsOk() && err.Message() == "keepalive watchdog timeout" && ms > 150.0 && ms < 2000.0);
  }
  mute.join();
  if (accepted >= 0) close(accepted);
  close(lfd);
}

// Both HTTP/2 implementations against each other with messages far beyond the flow-control windows:
// the client (GrpcChannel) sends 40 MB through the echo server's 4 MiB stream window (csrc/grpc_server.h
// replenishes it at the half-way mark) and receives 40 MB back through its own; then four 8 MB calls
// share the connection.  The echoed request parses as a response (same field numbers up to 6).
extern "C" {
struct tb200_grpc_echo_server;
int tb200_grpc_echo_server_start(const char* host, int* port, tb200_grpc_echo_server** out);
int tb200_grpc_echo_server_stop(tb200_grpc_echo_server* s);
}
static void TestLargeMessages() {
  int port = 0;
  tb200_grpc_echo_server* srv = nullptr;
  CHECK(tb200_grpc_echo_server_start("127.0.0.1", &port, &srv) == 0);
  {
    std::unique_ptr<tc::InferenceServerGrpcClient> client;
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, "127.0.0.1:" + std::to_string(port), false, false, tc::SslOptions(),
                                                   tc::KeepAliveOptions(), false));
    std::vector<int32_t> big(10 * 1000 * 1000);
    for (size_t i = 0; i < big.size(); ++i) big[i] = static_cast<int32_t>(i);
    tc::InferInput* in;
    tc::InferInput::Create(&in, "INPUT0", {static_cast<int64_t>(big.size())}, "INT32");
    in->AppendRaw(reinterpret_cast<uint8_t*>(big.data()), big.size() * 4);
    tc::InferOptions opt("echo");
    opt.request_id_ = "forty-megabytes";
    tc::InferResult* r = nullptr;
    CHECK_OK(client->Infer(&r, opt, {in}));
    std::string id;
    std::vector<int64_t> shape;
    if (r != nullptr) {
      CHECK_OK(r->Id(&id));
      CHECK_OK(r->Shape("INPUT0", &shape));  // the request's input metadata came back as output metadata
      CHECK(id == "forty-megabytes" && shape == std::vector<int64_t>({static_cast<int64_t>(big.size())}));
    }
    delete r;
    std::mutex mu;
    std::condition_variable cv;
    int done = 0, good = 0;
    tc::InferInput* mid;
    tc::InferInput::Create(&mid, "INPUT0", {2000000}, "INT32");
    mid->AppendRaw(reinterpret_cast<uint8_t*>(big.data()), 8000000);
    for (int i = 0; i < 4; ++i) {
      CHECK_OK(client->AsyncInfer(
          [&](tc::InferResult* res) {
            std::lock_guard<std::mutex> lk(mu);
            std::vector<int64_t> s;
            if (res->RequestStatus().IsOk() && res->Shape("INPUT0", &s).IsOk() && s == std::vector<int64_t>({2000000})) ++good;
            ++done;
            delete res;
            cv.notify_all();
          },
          opt, {mid}));
    }
    {
      std::unique_lock<std::mutex> lk(mu);
      CHECK(cv.wait_for(lk, std::chrono::seconds(60), [&] { return done == 4; }));
    }
    CHECK(good == 4);
    delete in;
    delete mid;
  }
  tb200_grpc_echo_server_stop(srv);
}

int main(int argc, char** argv) {
  if (argc > 2 && std::string(argv[1]) == "--roundtrip") return RoundTrip(argv[2]);
  if (argc > 1 && std::string(argv[1]) == "--requests") {
    PrintRequests();
    return 0;
  }
  TestCodec();
  TestResultDecoding();
  {
    std::unique_ptr<tc::InferenceServerGrpcClient> client;
    CHECK(!tc::InferenceServerGrpcClient::Create(&client, "localhost:1", false, true).IsOk());  // no TLS
    CHECK_OK(tc::InferenceServerGrpcClient::Create(&client, "127.0.0.1:1"));
    bool live = true;
    CHECK(!client->IsServerLive(&live).IsOk() && !live);  // nothing listens there
  }
  TestReconnect();
  TestKeepAlive();
  TestLargeMessages();
  if (argc > 2 && std::string(argv[2]) == "slow") TestSlowServer(argv[1]);
  else if (argc > 2 && std::string(argv[2]) == "compress-gpu") TestCompression(argv[1], true);
  else if (argc > 2 && std::string(argv[2]) == "compress-nogpu") TestCompression(argv[1], false);
  else if (argc > 2 && std::string(argv[2]) == "compressed-responses") TestCompression(argv[1], false, true);
  else if (argc > 1) TestLoopback(argv[1]);
  if (g_failures == 0) {
    std::cout << "PASS" << (argc > 1 ? " (offline + loopback)" : " (offline)") << std::endl;
    return 0;
  }
  std::cout << "FAIL: " << g_failures << " check(s)" << std::endl;
  return 1;
}
