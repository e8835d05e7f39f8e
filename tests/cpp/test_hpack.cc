// HPACK decoder of client_b200/csrc/h2.h against a real encoder: every line of stdin is one header
// block (hex) produced by ONE stateful libnghttp2 deflater (tests/test_hpack.py); they are decoded
// by ONE HpackDecoder (dynamic table, Huffman strings) and printed as hex "name value" pairs.
#include <iostream>
#include <string>
#include <vector>

#include "../../client_b200/csrc/h2.h"

static std::string Hex(const std::string& s) {
  static const char* d = "0123456789abcdef";
  std::string out;
  for (unsigned char c : s) {
    out.push_back(d[c >> 4]);
    out.push_back(d[c & 15]);
  }
  return out.empty() ? "-" : out;
}

int main() {
  tb200::h2::HpackDecoder decoder;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::string block;
    for (size_t i = 0; i + 1 < line.size(); i += 2) block.push_back(static_cast<char>(std::stoi(line.substr(i, 2), nullptr, 16)));
    std::vector<tb200::h2::HpackDecoder::Field> fields;
    if (!decoder.Decode(reinterpret_cast<const uint8_t*>(block.data()), block.size(), &fields)) {
      std::cout << "ERROR" << std::endl;
      continue;
    }
    for (const auto& f : fields) std::cout << Hex(f.first) << " " << Hex(f.second) << "\n";
    std::cout << "---" << std::endl;
  }
  return 0;
}
