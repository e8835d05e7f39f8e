"""Generate client_b200/cpp/grpc_service.pb.h: C++ message classes for the KServe-v2 / Triton gRPC
schema, from the table in client_b200/grpc/_proto.py (no protoc / libprotobuf in this image).

The classes offer the subset of the protoc C++ API the reference's gRPC client and examples use
(src/c++/library/grpc_client.cc, src/c++/examples/simple_grpc_*.cc): `x()`, `set_x()`,
`mutable_x()`, `add_x()`, `x_size()`, `has_x()`, `clear_x()`, maps as std::map, oneof `*_case()`,
`SerializeAsString` / `ParseFromString` / `DebugString`.  Run:  python scripts/gen_pb_cpp.py
(tests/test_cc_grpc_client.py checks the committed header is what this script prints)."""

import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from client_b200.grpc import _proto  # noqa: E402

VARINT = {"bool": "bool", "int32": "int32_t", "int64": "int64_t", "uint32": "uint32_t", "uint64": "uint64_t"}
SIGNED = {"int32", "int64"}
FIXED = {"float": ("float", "kFixed32", "fixed32", "float_bits", "bits_float"), "double": ("double", "kFixed64", "fixed64", "double_bits", "bits_double")}


def camel(name):
    return "".join(p.capitalize() for p in name.split("_"))


class Field:
    def __init__(self, owner, name, number, ftype, label):
        self.owner, self.name, self.number, self.label = owner, name, number, label
        self.oneof = label.split(":", 1)[1] if label.startswith("oneof:") else None
        self.repeated = label == "rep"
        self.map = label == "map" or label.startswith("map:")
        self.key = None
        if self.map:  # map<K, V>: K string or an integer type; the field's type is V
            self.key = Field(owner, "key", 1, label.split(":", 1)[1] if ":" in label else "string", "")
        self.enum = None
        self.msg = None
        top = owner.split(".")[0]
        if ftype.startswith("enum:"):
            e = ftype[5:]
            self.enum = ((top + e) if e.startswith(".") else e).replace(".", "_")
            self.kind = "varint"
            self.ctype = self.enum
        elif ftype in VARINT:
            self.kind, self.ctype = "varint", VARINT[ftype]
        elif ftype in FIXED:
            self.kind, self.ctype = "fixed", FIXED[ftype][0]
        elif ftype in ("string", "bytes"):
            self.kind, self.ctype = "string", "std::string"
        else:
            self.kind = "message"
            self.msg = (top + "." + ftype[1:]) if ftype.startswith(".") else ftype
            self.ctype = self.msg.replace(".", "_")
        self.ftype = ftype
        self.m = name + "_"

    # value -> uint64 for a varint
    def to_u64(self, v):
        if self.ftype in SIGNED or self.enum:
            return "static_cast<uint64_t>(static_cast<int64_t>(%s))" % v
        return "static_cast<uint64_t>(%s)" % v

    def from_u64(self, v):
        if self.ftype == "bool":
            return "(%s != 0)" % v
        return "static_cast<%s>(%s)" % (self.ctype, v)

    def elem_type(self):  # element type inside std::vector
        return "uint8_t" if self.ftype == "bool" else self.ctype


def collect():
    schemas = dict(_proto.MODEL_CONFIG_SCHEMA)
    schemas.update(_proto.SCHEMA)
    msgs = {name: [Field(name, *f) for f in fields] for name, fields in schemas.items()}
    order, seen = [], set()

    def visit(name):
        if name in seen:
            return
        seen.add(name)
        for child in msgs:  # nested types first
            if child.startswith(name + ".") :
                visit(child)
        for f in msgs[name]:
            if f.msg:
                visit(f.msg)
        order.append(name)

    for name in msgs:
        visit(name)
    return msgs, order


def emit_enum(out, qualified, values):
    cname = qualified.replace(".", "_")
    out.append("enum %s : int {" % cname)
    for n, v in values:
        out.append("  %s = %d," % (n if "." not in qualified else cname + "_" + n, v))
    out.append("};")
    out.append("inline const char* %s_Name(int v) {" % cname)
    out.append("  switch (v) {")
    for n, v in values:
        out.append('    case %d: return "%s";' % (v, n))
    out.append('    default: return "";')
    out.append("  }\n}")


def map_types(f):
    """(C++ key type, C++ value type) of a map field."""
    kt = "std::string" if f.key.kind == "string" else f.key.ctype
    return kt, f.ctype


def put_scalar(f, field_no, v, out="out"):
    """statement(s) appending one scalar / string / message value `v` as field `field_no`"""
    if f.kind == "message":
        return "{ std::string sub_; %s.AppendTo(&sub_); put_bytes(%s, %d, sub_); }" % (v, out, field_no)
    if f.kind == "string":
        return "put_bytes(%s, %d, %s);" % (out, field_no, v)
    if f.kind == "varint":
        return "{ put_tag(%s, %d, kVarint); put_varint(%s, %s); }" % (out, field_no, out, f.to_u64(v))
    _, wt, rd, tobits, _ = FIXED[f.ftype]
    return "{ put_tag(%s, %d, %s); put_%s(%s, %s(%s)); }" % (out, field_no, wt, rd, out, tobits, v)


def read_scalar(f, reader, target, wt="ewt"):
    """expression statement reading one value of f from Reader `reader` into lvalue `target`
    (caller checked the field number); returns code that evaluates to true when consumed"""
    if f.kind == "message":
        return "(%s == kBytes && %s.bytes(&ed, &elen) && ((r.ok = %s.MergeFrom(ed, elen) && r.ok), true))" % (wt, reader, target)
    if f.kind == "string":
        return "(%s == kBytes && %s.bytes(&ed, &elen) && (%s.assign(reinterpret_cast<const char*>(ed), elen), true))" % (wt, reader, target)
    if f.kind == "varint":
        return "(%s == kVarint && ((%s = %s), true))" % (wt, target, f.from_u64("%s.varint()" % reader))
    _, wtn, rd, _, frombits = FIXED[f.ftype]
    return "(%s == %s && ((%s = %s(%s.%s())), true))" % (wt, wtn, target, frombits, reader, rd)


def generate():
    msgs, order = collect()
    out = []
    w = out.append
    w("// grpc_service.pb.h -- GENERATED by scripts/gen_pb_cpp.py from client_b200/grpc/_proto.py; do not edit.")
    w("// Message classes of inference.GRPCInferenceService (schema: src/rust/triton-client/proto/")
    w("// grpc_service.proto, model_config.proto subset) with the protoc-style accessors the reference's")
    w("// C++ gRPC client uses.  Runtime: pb.h.")
    w("#ifndef TB200_CPP_GRPC_SERVICE_PB_H_")
    w("#define TB200_CPP_GRPC_SERVICE_PB_H_")
    w("")
    w("#include <cstdint>\n#include <map>\n#include <string>\n#include <utility>\n#include <vector>")
    w("")
    w('#include "pb.h"')
    w("")
    w("namespace inference {")
    w("")
    for qualified, values in _proto.MODEL_CONFIG_ENUMS.items():
        emit_enum(out, qualified, values)
    w("")
    for name in order:
        fields = sorted(msgs[name], key=lambda f: f.number)
        cname = name.replace(".", "_")
        oneofs = {}
        for f in fields:
            if f.oneof:
                oneofs.setdefault(f.oneof, []).append(f)
        w("class %s : public tb200::pb::Message {" % cname)
        w(" public:")
        for child in msgs:
            if child.startswith(name + ".") and "." not in child[len(name) + 1:]:
                w("  using %s = %s;" % (child[len(name) + 1:], child.replace(".", "_")))
        for qualified, values in _proto.MODEL_CONFIG_ENUMS.items():  # enums nested in this message
            if qualified.startswith(name + ".") and "." not in qualified[len(name) + 1:]:
                short, ecname = qualified[len(name) + 1:], qualified.replace(".", "_")
                w("  using %s = %s;" % (short, ecname))
                for n, v in values:
                    w("  static constexpr %s %s = %s_%s;" % (short, n, ecname, n))
        for group, members in oneofs.items():
            w("  enum %sCase { %s %s_NOT_SET = 0 };" % (camel(group), " ".join("k%s = %d," % (camel(m.name), m.number) for m in members), group.upper()))
            w("  %sCase %s_case() const { return static_cast<%sCase>(%s_case_); }" % (camel(group), group, camel(group), group))
        # ---- accessors
        for f in fields:
            n, m, ct = f.name, f.m, f.ctype
            if f.map:
                kt, vt = map_types(f)
                w("  const std::map<%s, %s>& %s() const { return %s; }" % (kt, vt, n, m))
                w("  std::map<%s, %s>* mutable_%s() { return &%s; }" % (kt, vt, n, m))
                w("  int %s_size() const { return static_cast<int>(%s.size()); }" % (n, m))
                w("  void clear_%s() { %s.clear(); }" % (n, m))
            elif f.repeated:
                et = f.elem_type()
                w("  int %s_size() const { return static_cast<int>(%s.size()); }" % (n, m))
                w("  const std::vector<%s>& %s() const { return %s; }" % (et, n, m))
                w("  std::vector<%s>* mutable_%s() { return &%s; }" % (et, n, m))
                w("  void clear_%s() { %s.clear(); }" % (n, m))
                if f.kind == "message":
                    w("  const %s& %s(int i) const { return %s[static_cast<size_t>(i)]; }" % (ct, n, m))
                    w("  %s* mutable_%s(int i) { return &%s[static_cast<size_t>(i)]; }" % (ct, n, m))
                    w("  %s* add_%s() {\n    %s.emplace_back();\n    return &%s.back();\n  }" % (ct, n, m, m))
                elif f.kind == "string":
                    w("  const std::string& %s(int i) const { return %s[static_cast<size_t>(i)]; }" % (n, m))
                    w("  std::string* mutable_%s(int i) { return &%s[static_cast<size_t>(i)]; }" % (n, m))
                    w("  void add_%s(const std::string& v) { %s.push_back(v); }" % (n, m))
                    w("  void add_%s(const char* v) { %s.emplace_back(v); }" % (n, m))
                    w("  void add_%s(const void* v, size_t n) { %s.emplace_back(static_cast<const char*>(v), n); }" % (n, m))
                    w("  std::string* add_%s() {\n    %s.emplace_back();\n    return &%s.back();\n  }" % (n, m, m))
                else:
                    w("  %s %s(int i) const { return static_cast<%s>(%s[static_cast<size_t>(i)]); }" % (ct, n, ct, m))
                    w("  void add_%s(%s v) { %s.push_back(static_cast<%s>(v)); }" % (n, ct, m, et))
            elif f.kind == "message" and f.oneof:
                w("  bool has_%s() const { return %s_case_ == %d; }" % (n, f.oneof, f.number))
                w("  const %s& %s() const {\n    static const tb200::pb::Box<%s> kNone;\n    return %s_case_ == %d ? %s.get() : kNone.get();\n  }" % (ct, n, ct, f.oneof, f.number, m))
                w("  %s* mutable_%s() {\n    %s_case_ = %d;\n    return %s.mut();\n  }" % (ct, n, f.oneof, f.number, m))
                w("  void clear_%s() {\n    if (%s_case_ == %d) %s_case_ = 0;\n    %s.reset();\n  }" % (n, f.oneof, f.number, f.oneof, m))
            elif f.kind == "message":
                w("  bool has_%s() const { return %s.has(); }" % (n, m))
                w("  const %s& %s() const { return %s.get(); }" % (ct, n, m))
                w("  %s* mutable_%s() { return %s.mut(); }" % (ct, n, m))
                w("  void clear_%s() { %s.reset(); }" % (n, m))
            else:
                setcase = ("    %s_case_ = %d;\n" % (f.oneof, f.number)) if f.oneof else ""
                if f.oneof:
                    w("  bool has_%s() const { return %s_case_ == %d; }" % (n, f.oneof, f.number))
                if f.kind == "string":
                    if f.oneof:
                        w("  const std::string& %s() const {\n    static const std::string kEmpty;\n    return %s_case_ == %d ? %s : kEmpty;\n  }" % (n, f.oneof, f.number, m))
                    else:
                        w("  const std::string& %s() const { return %s; }" % (n, m))
                    w("  void set_%s(const std::string& v) {\n%s    %s = v;\n  }" % (n, setcase, m))
                    w("  void set_%s(const char* v) {\n%s    %s = v;\n  }" % (n, setcase, m))
                    w("  void set_%s(const void* v, size_t n) {\n%s    %s.assign(static_cast<const char*>(v), n);\n  }" % (n, setcase, m))
                    w("  std::string* mutable_%s() {\n%s    return &%s;\n  }" % (n, setcase, m))
                else:
                    zero = "false" if f.ftype == "bool" else ("static_cast<%s>(0)" % ct)
                    if f.oneof:
                        w("  %s %s() const { return %s_case_ == %d ? %s : %s; }" % (ct, n, f.oneof, f.number, m, zero))
                    else:
                        w("  %s %s() const { return %s; }" % (ct, n, m))
                    w("  void set_%s(%s v) {\n%s    %s = v;\n  }" % (n, ct, setcase, m))
                if f.oneof:
                    w("  void clear_%s() {\n    if (%s_case_ == %d) %s_case_ = 0;\n  }" % (n, f.oneof, f.number, f.oneof))
                elif f.kind == "string":
                    w("  void clear_%s() { %s.clear(); }" % (n, m))
                else:
                    w("  void clear_%s() { %s = %s; }" % (n, m, "false" if f.ftype == "bool" else "static_cast<%s>(0)" % ct))
        # ---- Clear
        w("")
        w("  void Clear() override {")
        for f in fields:
            if f.map or f.repeated or f.kind == "string":
                w("    %s.clear();" % f.m)
            elif f.kind == "message":
                w("    %s.reset();" % f.m)
            else:
                w("    %s = %s;" % (f.m, "false" if f.ftype == "bool" else "static_cast<%s>(0)" % f.ctype))
        for group in oneofs:
            w("    %s_case_ = 0;" % group)
        w("  }")
        # ---- serialise
        w("  void AppendTo(std::string* out) const override {")
        w("    using namespace tb200::pb;")
        w("    (void)out;")
        for f in fields:
            N, m = f.number, f.m
            if f.map:
                w("    for (const auto& kv : %s) {" % m)
                w("      std::string entry;\n      %s\n      %s\n      put_bytes(out, %d, entry);\n    }" % (
                    put_scalar(f.key, 1, "kv.first", "&entry"), put_scalar(f, 2, "kv.second", "&entry"), N))
            elif f.repeated:
                if f.kind == "message":
                    w("    for (const auto& v : %s) {\n      std::string sub;\n      v.AppendTo(&sub);\n      put_bytes(out, %d, sub);\n    }" % (m, N))
                elif f.kind == "string":
                    w("    for (const auto& v : %s) put_bytes(out, %d, v);" % (m, N))
                elif f.kind == "varint":
                    w("    if (!%s.empty()) {\n      std::string sub;\n      for (auto v : %s) put_varint(&sub, %s);\n      put_bytes(out, %d, sub);\n    }" % (m, m, f.to_u64("v"), N))
                else:
                    _, _, rd, tobits, _ = FIXED[f.ftype]
                    w("    if (!%s.empty()) {\n      std::string sub;\n      for (auto v : %s) put_%s(&sub, %s(v));\n      put_bytes(out, %d, sub);\n    }" % (m, m, rd, tobits, N))
            elif f.kind == "message":
                cond = ("%s_case_ == %d" % (f.oneof, N)) if f.oneof else ("%s.has()" % m)
                w("    if (%s) {\n      std::string sub;\n      %s.get().AppendTo(&sub);\n      put_bytes(out, %d, sub);\n    }" % (cond, m, N))
            else:
                if f.oneof:
                    cond = "%s_case_ == %d" % (f.oneof, N)
                elif f.kind == "string":
                    cond = "!%s.empty()" % m
                elif f.kind == "fixed":
                    cond = "%s(%s) != 0" % (FIXED[f.ftype][3], m)
                else:
                    cond = "%s != 0" % m if f.ftype != "bool" else m
                if f.kind == "string":
                    w("    if (%s) put_bytes(out, %d, %s);" % (cond, N, m))
                elif f.kind == "varint":
                    w("    if (%s) {\n      put_tag(out, %d, kVarint);\n      put_varint(out, %s);\n    }" % (cond, N, f.to_u64(m)))
                else:
                    _, wt, rd, tobits, _ = FIXED[f.ftype]
                    w("    if (%s) {\n      put_tag(out, %d, %s);\n      put_%s(out, %s(%s));\n    }" % (cond, N, wt, rd, tobits, m))
        w("  }")
        # ---- parse
        w("  bool MergeFrom(const uint8_t* data, size_t n) override {")
        w("    using namespace tb200::pb;")
        w("    Reader r(data, n);")
        w("    uint32_t field = 0, wt = 0;")
        w("    while (r.tag(&field, &wt)) {")
        w("      switch (field) {")
        for f in fields:
            N, m = f.number, f.m
            w("        case %d: {" % N)
            if f.map or f.repeated or f.kind in ("message", "string"):
                w("          const uint8_t* d = nullptr;\n          size_t len = 0;")
            if f.map:
                kt, vt = map_types(f)
                kinit = "" if f.key.kind == "string" else " = 0"
                vinit = "" if f.kind in ("message", "string") else (" = false" if f.ftype == "bool" else " = static_cast<%s>(0)" % vt)
                w("          if (wt == kBytes && r.bytes(&d, &len)) {")
                w("            Reader e(d, len);\n            %s key%s;\n            %s value%s;\n            uint32_t ef = 0, ewt = 0;" % (kt, kinit, vt, vinit))
                w("            while (e.tag(&ef, &ewt)) {\n              const uint8_t* ed = nullptr;\n              size_t elen = 0;\n              (void)ed;\n              (void)elen;")
                w("              if (ef == 1 && %s) {" % read_scalar(f.key, "e", "key"))
                w("              } else if (ef == 2 && %s) {" % read_scalar(f, "e", "value"))
                w("              } else {\n                e.skip(ewt);\n              }\n            }")
                w("            if (!e.ok) r.ok = false;\n            %s[key] = std::move(value);\n          } else {\n            r.skip(wt);\n          }" % m)
            elif f.repeated and f.kind == "message":
                w("          if (wt == kBytes && r.bytes(&d, &len)) {\n            %s.emplace_back();\n            if (!%s.back().MergeFrom(d, len)) r.ok = false;\n          } else {\n            r.skip(wt);\n          }" % (m, m))
            elif f.repeated and f.kind == "string":
                w("          if (wt == kBytes && r.bytes(&d, &len)) %s.emplace_back(reinterpret_cast<const char*>(d), len);\n          else r.skip(wt);" % m)
            elif f.repeated and f.kind == "varint":
                et = f.elem_type()
                w("          if (wt == kBytes && r.bytes(&d, &len)) {\n            Reader p(d, len);\n            while (!p.done()) {\n              const uint64_t v = p.varint();\n              if (p.ok) %s.push_back(static_cast<%s>(%s));\n            }\n            if (!p.ok) r.ok = false;" % (m, et, f.from_u64("v")))
                w("          } else if (wt == kVarint) {\n            const uint64_t v = r.varint();\n            %s.push_back(static_cast<%s>(%s));\n          } else {\n            r.skip(wt);\n          }" % (m, et, f.from_u64("v")))
            elif f.repeated:
                _, wtn, rd, _, frombits = FIXED[f.ftype]
                w("          if (wt == kBytes && r.bytes(&d, &len)) {\n            Reader p(d, len);\n            while (!p.done()) {\n              const auto v = p.%s();\n              if (p.ok) %s.push_back(%s(v));\n            }\n            if (!p.ok) r.ok = false;" % (rd, m, frombits))
                w("          } else if (wt == %s) {\n            %s.push_back(%s(r.%s()));\n          } else {\n            r.skip(wt);\n          }" % (wtn, m, frombits, rd))
            elif f.kind == "message":
                setcase = ("            %s_case_ = %d;\n" % (f.oneof, N)) if f.oneof else ""
                w("          if (wt == kBytes && r.bytes(&d, &len)) {\n%s            if (!%s.mut()->MergeFrom(d, len)) r.ok = false;\n          } else {\n            r.skip(wt);\n          }" % (setcase, m))
            else:
                setcase = ("            %s_case_ = %d;\n" % (f.oneof, N)) if f.oneof else ""
                if f.kind == "string":
                    w("          if (wt == kBytes && r.bytes(&d, &len)) {\n%s            %s.assign(reinterpret_cast<const char*>(d), len);\n          } else {\n            r.skip(wt);\n          }" % (setcase, m))
                elif f.kind == "varint":
                    w("          if (wt == kVarint) {\n            const uint64_t v = r.varint();\n%s            %s = %s;\n          } else {\n            r.skip(wt);\n          }" % (setcase, m, f.from_u64("v")))
                else:
                    _, wtn, rd, _, frombits = FIXED[f.ftype]
                    w("          if (wt == %s) {\n%s            %s = %s(r.%s());\n          } else {\n            r.skip(wt);\n          }" % (wtn, setcase, m, frombits, rd))
            w("          break;\n        }")
        w("        default:\n          r.skip(wt);\n      }\n    }\n    return r.ok;\n  }")
        # ---- text
        w("  void PrintTo(std::string* out, int indent) const override {")
        w("    using namespace tb200::pb;")
        w("    (void)out;\n    (void)indent;")

        def scalar_text(f, v):
            if f.kind == "string":
                return "text_escaped(out, %s);" % v
            if f.enum:  # an unknown value prints as its number, like libprotobuf's text format
                return ("{ const char* nm_ = %s_Name(static_cast<int>(%s)); if (*nm_) out->append(nm_); else out->append(std::to_string(static_cast<int>(%s))); }"
                        % (f.enum, v, v))
            if f.ftype == "bool":
                return 'out->append(%s ? "true" : "false");' % v
            if f.kind == "fixed":
                return "text_float(out, static_cast<double>(%s), %s);" % (v, "true" if f.ftype == "float" else "false")
            if f.ftype in ("int32", "int64"):
                return "out->append(std::to_string(static_cast<long long>(%s)));" % v
            return "out->append(std::to_string(static_cast<unsigned long long>(%s)));" % v

        for f in fields:
            n, m = f.name, f.m
            if f.map:
                w("    for (const auto& kv : %s) {" % m)
                w('      text_indent(out, indent);\n      out->append("%s {\\n");' % n)
                def nondefault(ff, v):
                    return ("!%s.empty()" % v) if ff.kind == "string" else ("%s != 0" % v if ff.ftype != "bool" else v)

                # default-valued keys / scalar values of an entry are not printed (proto3 text format)
                w('      if (%s) {\n        text_indent(out, indent + 2);\n        out->append("key: ");\n        %s\n        out->append("\\n");\n      }' % (nondefault(f.key, "kv.first"), scalar_text(f.key, "kv.first")))
                if f.kind == "message":
                    w('      text_indent(out, indent + 2);\n      out->append("value {\\n");\n      kv.second.PrintTo(out, indent + 4);\n      text_indent(out, indent + 2);\n      out->append("}\\n");')
                else:
                    w('      if (%s) {\n        text_indent(out, indent + 2);\n        out->append("value: ");\n        %s\n        out->append("\\n");\n      }' % (nondefault(f, "kv.second"), scalar_text(f, "kv.second")))
                w('      text_indent(out, indent);\n      out->append("}\\n");\n    }')
            elif f.kind == "message":
                if f.repeated:
                    w("    for (const auto& v : %s) {" % m)
                    w('      text_indent(out, indent);\n      out->append("%s {\\n");\n      v.PrintTo(out, indent + 2);\n      text_indent(out, indent);\n      out->append("}\\n");\n    }' % n)
                else:
                    w("    if (%s) {" % (("%s_case_ == %d" % (f.oneof, f.number)) if f.oneof else ("%s.has()" % m)))
                    w('      text_indent(out, indent);\n      out->append("%s {\\n");\n      %s.get().PrintTo(out, indent + 2);\n      text_indent(out, indent);\n      out->append("}\\n");\n    }' % (n, m))
            elif f.repeated:
                w("    for (const auto& v : %s) {" % m)
                w('      text_indent(out, indent);\n      out->append("%s: ");\n      %s\n      out->append("\\n");\n    }' % (n, scalar_text(f, "v")))
            else:
                if f.oneof:
                    cond = "%s_case_ == %d" % (f.oneof, f.number)
                elif f.kind == "string":
                    cond = "!%s.empty()" % m
                elif f.kind == "fixed":
                    cond = "%s(%s) != 0" % (FIXED[f.ftype][3], m)
                else:
                    cond = m if f.ftype == "bool" else "%s != 0" % m
                w("    if (%s) {" % cond)
                w('      text_indent(out, indent);\n      out->append("%s: ");\n      %s\n      out->append("\\n");\n    }' % (n, scalar_text(f, m)))
        w("  }")
        # ---- members
        w("")
        w(" private:")
        for f in fields:
            if f.map:
                w("  std::map<%s, %s> %s;" % (map_types(f) + (f.m,)))
            elif f.repeated:
                w("  std::vector<%s> %s;" % (f.elem_type(), f.m))
            elif f.kind == "message":
                w("  tb200::pb::Box<%s> %s;" % (f.ctype, f.m))
            elif f.kind == "string":
                w("  std::string %s;" % f.m)
            else:
                w("  %s %s = %s;" % (f.ctype, f.m, "false" if f.ftype == "bool" else "static_cast<%s>(0)" % f.ctype))
        for group in oneofs:
            w("  int %s_case_ = 0;" % group)
        w("};")
        w("")
    w("}  // namespace inference")
    w("")
    w("#endif  // TB200_CPP_GRPC_SERVICE_PB_H_")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = generate()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "client_b200", "cpp", "grpc_service.pb.h")
    if "--check" in sys.argv:
        sys.exit(0 if open(path).read() == text else 1)
    with open(path, "w") as f:
        f.write(text)
    print("wrote %s (%d lines)" % (os.path.normpath(path), text.count("\n")))
