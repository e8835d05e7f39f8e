// CUDA-shared-memory flow of the C++ front end on a GPU: the flow of
// src/c++/examples/simple_http_cudashm_client.cc:120-290 with the AppendRaw scatter list
// gathered straight into the IPC region and the outputs validated on the device.
// Usage: test_cc_cudashm host:port fill_dump_path
#include <cstdio>
#include <cstring>
#include <iostream>

#include "http_client.h"

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

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::unique_ptr<tc::InferenceServerHttpClient> client;
  CHECK_OK(tc::InferenceServerHttpClient::Create(&client, argv[1]));
  CHECK_OK(client->UnregisterCudaSharedMemory());

  std::unique_ptr<tc::CudaRegion> in, out;
  CHECK_OK(tc::CudaRegion::Create(&in, "cc_input_data", 128, 0));
  CHECK_OK(tc::CudaRegion::Create(&out, "cc_output_data", 128, 0));

  // INPUT0 arrives as two borrowed halves, INPUT1 as one buffer: one gathered transfer each
  int32_t a[16], b[16];
  for (int i = 0; i < 16; ++i) {
    a[i] = i;
    b[i] = 1;
  }
  tc::InferInput *in0 = nullptr, *in1 = nullptr;
  tc::InferInput::Create(&in0, "INPUT0", {1, 16}, "INT32");
  tc::InferInput::Create(&in1, "INPUT1", {1, 16}, "INT32");
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a), 32);
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a) + 32, 32);
  in1->AppendRaw(reinterpret_cast<uint8_t*>(b), 64);
  CHECK_OK(in->SetFromInput(*in0, 0));
  CHECK_OK(in->SetFromInput(*in1, 64));
  int32_t back[32];
  CHECK_OK(in->Read(0, back, 128));
  CHECK(memcmp(back, a, 64) == 0 && memcmp(back + 16, b, 64) == 0);

  CHECK_OK(in->Register(client.get()));
  CHECK_OK(out->Register(client.get()));
  std::string status;
  CHECK_OK(client->CudaSharedMemoryStatus(&status));
  CHECK(status.find("cc_input_data") != std::string::npos && status.find("cc_output_data") != std::string::npos);

  // the request only names the regions
  in0->SetSharedMemory("cc_input_data", 64, 0);
  in1->SetSharedMemory("cc_input_data", 64, 64);
  tc::InferRequestedOutput *o0 = nullptr, *o1 = nullptr;
  tc::InferRequestedOutput::Create(&o0, "OUTPUT0");
  tc::InferRequestedOutput::Create(&o1, "OUTPUT1");
  o0->SetSharedMemory("cc_output_data", 64, 0);
  o1->SetSharedMemory("cc_output_data", 64, 64);
  tc::InferOptions options("simple");
  tc::InferResult* result = nullptr;
  CHECK_OK(client->Infer(&result, options, {in0, in1}, {o0, o1}));
  if (result != nullptr) {
    std::vector<int64_t> shape;
    CHECK_OK(result->Shape("OUTPUT0", &shape));
    CHECK((shape == std::vector<int64_t>{1, 16}));
    delete result;
  }
  // validated on the device, then once more on the host
  uint64_t mismatches = 99;
  CHECK_OK(out->CheckAddSub(0, 64, *in, 0, 64, 64, &mismatches));
  CHECK(mismatches == 0);
  CHECK_OK(out->Read(0, back, 128));
  for (int i = 0; i < 16; ++i) CHECK(back[i] == a[i] + b[i] && back[16 + i] == a[i] - b[i]);
  // a corrupted output is caught
  int32_t wrong = 12345;
  CHECK_OK(out->Write(8, &wrong, 4));
  CHECK_OK(out->CheckAddSub(0, 64, *in, 0, 64, 64, &mismatches));
  CHECK(mismatches == 1);

  // synthetic input generated in place (compared with the oracle by the pytest wrapper)
  std::unique_ptr<tc::CudaRegion> big;
  CHECK_OK(tc::CudaRegion::Create(&big, "cc_fill", 602112, 0));
  CHECK_OK(big->FillRandom(0, "FP32", 602112, 7, 3));
  std::vector<uint8_t> host(602112);
  CHECK_OK(big->Read(0, host.data(), host.size()));
  FILE* f = fopen(argv[2], "wb");
  CHECK(f != nullptr && fwrite(host.data(), 1, host.size(), f) == host.size());
  if (f) fclose(f);
  CHECK(!big->FillRandom(0, "BYTES", 16, 0, 0).IsOk());
  CHECK(!big->FillRandom(602112, "FP32", 16, 0, 0).IsOk());

  CHECK_OK(client->UnregisterCudaSharedMemory("cc_input_data"));
  CHECK_OK(client->UnregisterCudaSharedMemory());
  delete in0;
  delete in1;
  delete o0;
  delete o1;
  std::cout << (g_failures == 0 ? "PASS" : "FAIL") << std::endl;
  return g_failures == 0 ? 0 : 1;
}
