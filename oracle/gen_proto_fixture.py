"""gen_proto_fixture.py -- TEST INFRASTRUCTURE (oracle/): parse the reference's copy of the wire
schema into a committed fixture.

The only schema files in the reference are the vendored .proto files of its Rust client,
/root/reference/src/rust/triton-client/proto/{grpc_service,model_config,health}.proto (the Python
and C++ clients get generated stubs from the `common` repository at build time).
This script reads them with a small proto3 parser (no protoc in this image) and writes

    tests/golden/proto_schema.json    every message (nested ones by qualified name) with its
                                      fields (name, number, type, label, oneof, map key/value),
                                      every enum with its values, every service with its rpcs

`tests/test_proto_parity.py` diffs that fixture against the descriptors client_b200/grpc/_proto.py
builds at import time, so a wrong field number or type in the hand-written table no longer
"agrees with itself".  With --emit-model-config it also prints the ModelConfig half of the table
(client_b200/grpc/_model_config_schema.py), which is schema data, not code: names, numbers, types.

Run where the reference tree is mounted:   python oracle/gen_proto_fixture.py [--emit-model-config]
"""

import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROTO_DIR = "/root/reference/src/rust/triton-client/proto"
FILES = ["grpc_service.proto", "model_config.proto", "health.proto"]
SCALARS = {"double", "float", "int32", "int64", "uint32", "uint64", "sint32", "sint64", "fixed32", "fixed64",
           "sfixed32", "sfixed64", "bool", "string", "bytes"}

_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[A-Za-z_][A-Za-z0-9_.]*|-?\d+|[{}()\[\]<>=;,]')


def tokenize(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return _TOKEN.findall(text)


class Parser:
    def __init__(self, tokens):
        self.t, self.i = tokens, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self, expect=None):
        tok = self.t[self.i]
        if expect is not None and tok != expect:
            raise ValueError("expected %r, got %r at token %d" % (expect, tok, self.i))
        self.i += 1
        return tok

    def skip_statement(self):
        depth = 0
        while True:
            tok = self.take()
            if tok == "{":
                depth += 1
            elif tok == "}":
                depth -= 1
                if depth == 0:
                    return
            elif tok == ";" and depth == 0:
                return

    def parse_file(self):
        out = {"package": "", "messages": {}, "enums": {}, "services": {}}
        while self.peek() is not None:
            tok = self.peek()
            if tok == "syntax":
                self.take(); self.take("="); out["syntax"] = self.take().strip('"'); self.take(";")
            elif tok == "package":
                self.take(); out["package"] = self.take(); self.take(";")
            elif tok == "message":
                self.parse_message("", out)
            elif tok == "enum":
                self.parse_enum("", out)
            elif tok == "service":
                self.parse_service(out)
            else:
                self.skip_statement()  # import / option
        return out

    def parse_enum(self, scope, out):
        self.take("enum")
        name = scope + self.take()
        self.take("{")
        values = []
        while self.peek() != "}":
            if self.peek() in ("option", "reserved"):
                self.skip_statement()
                continue
            vname = self.take(); self.take("=")
            values.append([vname, int(self.take())])
            if self.peek() == "[":
                while self.take() != "]":
                    pass
            self.take(";")
        self.take("}")
        out["enums"][name] = values

    def parse_field(self, msg, oneof):
        label = ""
        if self.peek() in ("repeated", "optional", "required"):
            label = self.take()
        ftype = self.take()
        field = {"label": "repeated" if label == "repeated" else ("optional" if label == "optional" else "")}
        if ftype == "map":
            self.take("<"); k = self.take(); self.take(","); v = self.take(); self.take(">")
            field.update(type="map", key=k, value=v)
        else:
            field["type"] = ftype
        field["name"] = self.take()
        self.take("=")
        field["number"] = int(self.take())
        if self.peek() == "[":
            while self.take() != "]":
                pass
        self.take(";")
        if oneof:
            field["oneof"] = oneof
        msg["fields"].append(field)

    def parse_message(self, scope, out):
        self.take("message")
        name = scope + self.take()
        msg = {"fields": []}
        out["messages"][name] = msg
        self.take("{")
        while self.peek() != "}":
            tok = self.peek()
            if tok == "message":
                self.parse_message(name + ".", out)
            elif tok == "enum":
                self.parse_enum(name + ".", out)
            elif tok == "oneof":
                self.take(); group = self.take(); self.take("{")
                while self.peek() != "}":
                    self.parse_field(msg, group)
                self.take("}")
            elif tok in ("option", "reserved", "extensions"):
                self.skip_statement()
            else:
                self.parse_field(msg, None)
        self.take("}")

    def parse_service(self, out):
        self.take("service")
        name = self.take()
        self.take("{")
        rpcs = {}
        while self.peek() != "}":
            if self.peek() == "option":
                self.skip_statement()
                continue
            self.take("rpc")
            rpc = self.take(); self.take("(")
            cs = self.peek() == "stream"
            if cs:
                self.take()
            req = self.take(); self.take(")"); self.take("returns"); self.take("(")
            ss = self.peek() == "stream"
            if ss:
                self.take()
            resp = self.take(); self.take(")")
            if self.peek() == "{":
                self.skip_statement()
            else:
                self.take(";")
            rpcs[rpc] = [req, resp, cs, ss]
        self.take("}")
        out["services"][name] = rpcs


def resolve(type_name, scope, known):
    """proto scoping: innermost enclosing scope outwards; returns the qualified name."""
    if type_name in SCALARS:
        return type_name
    parts = scope.split(".") if scope else []
    for depth in range(len(parts), -1, -1):
        cand = ".".join(parts[:depth] + [type_name])
        if cand in known:
            return cand
    raise ValueError("unresolved type %s in %s" % (type_name, scope))


def load():
    packages = {}
    for fn in FILES:
        with open(os.path.join(PROTO_DIR, fn)) as fh:
            parsed = Parser(tokenize(fh.read())).parse_file()
        pkg = packages.setdefault(parsed["package"], {"messages": {}, "enums": {}, "services": {}, "files": []})
        pkg["files"].append(fn)
        for k in ("messages", "enums", "services"):
            pkg[k].update(parsed[k])
    for pkg in packages.values():
        known = set(pkg["messages"]) | set(pkg["enums"])
        for mname, msg in pkg["messages"].items():
            for f in msg["fields"]:
                for key in ("type", "key", "value"):
                    if key in f and f[key] != "map":
                        f[key] = resolve(f[key], mname, known)
                        if key != "key":
                            f["kind"] = "scalar" if f[key] in SCALARS else ("enum" if f[key] in pkg["enums"] else "message")
    return packages


def to_table(pkg, names):
    """The compact table form of client_b200/grpc/_proto.py for the given top-level messages
    (and everything nested in / referenced by them): name -> [(field, number, type, label)]."""
    lines = ['"""ModelConfig half of the gRPC schema table -- GENERATED by oracle/gen_proto_fixture.py',
             "--emit-model-config from the reference's src/rust/triton-client/proto/model_config.proto",
             "(schema data: message / field names, numbers and types; tests/test_proto_parity.py diffs the",
             'descriptors built from it against the same file).  Do not edit."""', "",
             "# message -> [(field, number, type, label)]; types: scalar | qualified message | 'enum:<qualified>';",
             "# label: '' | 'rep' | 'map:<key type>' (type = value type) | 'oneof:<group>'",
             "MODEL_CONFIG_SCHEMA = {"]
    for mname in names:
        fields = []
        for f in pkg["messages"][mname]["fields"]:
            if f["type"] == "map":
                t = ("enum:" + f["value"]) if f["kind"] == "enum" else f["value"]
                label = "map:" + f["key"]
            else:
                t = ("enum:" + f["type"]) if f["kind"] == "enum" else f["type"]
                label = "rep" if f["label"] == "repeated" else ("oneof:" + f["oneof"] if "oneof" in f else "")
            fields.append('("%s", %d, "%s", "%s")' % (f["name"], f["number"], t, label))
        lines.append('    "%s": [%s],' % (mname, ", ".join(fields)))
    lines.append("}")
    lines.append("")
    lines.append("MODEL_CONFIG_ENUMS = {")
    for ename, values in pkg["enums"].items():
        lines.append('    "%s": [%s],' % (ename, ", ".join('("%s", %d)' % (n, v) for n, v in values)))
    lines.append("}")
    return "\n".join(lines) + "\n"


def main():
    packages = load()
    path = os.path.join(ROOT, "tests", "golden", "proto_schema.json")
    with open(path, "w") as fh:
        json.dump({"source": [os.path.join("src/rust/triton-client/proto", f) for f in FILES], "packages": packages}, fh, sort_keys=True, separators=(",", ":"))
        fh.write("\n")
    n_msg = sum(len(p["messages"]) for p in packages.values())
    print("wrote %s: %d messages, %d enums" % (path, n_msg, sum(len(p["enums"]) for p in packages.values())))
    if "--emit-model-config" in sys.argv:
        with open(os.path.join(PROTO_DIR, "model_config.proto")) as fh:
            own = Parser(tokenize(fh.read())).parse_file()
        inf = packages["inference"]
        out = os.path.join(ROOT, "client_b200", "grpc", "_model_config_schema.py")
        with open(out, "w") as fh:
            fh.write(to_table({"messages": {k: inf["messages"][k] for k in own["messages"]}, "enums": {k: inf["enums"][k] for k in own["enums"]}},
                              list(own["messages"])))
        print("wrote %s" % out)


if __name__ == "__main__":
    main()
