"""Generates tests/golden/proto/*.iirb (protobuf wire format of serialized StencilInstantiations) and
tests/golden/proto/schema.txt with the OFFICIAL protobuf runtime (google.protobuf) driven by the
reference's own schema files, read where they lie under /root/reference:

    dawn/src/dawn/IIR/proto/IIR/IIR.proto, dawn/src/dawn/AST/proto/AST/{statements,enums}.proto

There is no protoc in this image, so a small .proto reader below turns the three files into
FileDescriptorProtos; everything after that (JSON parsing, serialization) is google.protobuf.
Run in the build container only (needs /root/reference):  python tests/golden/make_proto_golden.py
"""
import glob
import os
import re
import sys

from google.protobuf import descriptor_pb2, descriptor_pool, json_format, message_factory
from google.protobuf import wrappers_pb2  # noqa: F401  (registers google/protobuf/wrappers.proto)

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/dawn/src/dawn"
FILES = [("AST/enums.proto", REF + "/AST/proto/AST/enums.proto"),
         ("AST/statements.proto", REF + "/AST/proto/AST/statements.proto"),
         ("IIR/IIR.proto", REF + "/IIR/proto/IIR/IIR.proto")]
SCALARS = {"int32": 5, "int64": 3, "bool": 8, "double": 1, "float": 2, "string": 9, "uint32": 13}
F = descriptor_pb2.FieldDescriptorProto


def tokens(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return re.findall(r"[A-Za-z_][A-Za-z_0-9.]*|\d+|\"[^\"]*\"|[{}=;<>,]", text)


class ProtoReader:
    def __init__(self, name, path):
        self.t = tokens(open(path).read())
        self.p = 0
        self.fd = descriptor_pb2.FileDescriptorProto(name=name, syntax="proto3")

    def peek(self):
        return self.t[self.p]

    def take(self, want=None):
        tok = self.t[self.p]
        self.p += 1
        assert want is None or tok == want, (want, tok, self.t[self.p - 5:self.p + 5])
        return tok

    def parse(self):
        while self.p < len(self.t):
            tok = self.take()
            if tok == "syntax":
                self.take("="), self.take(), self.take(";")
            elif tok == "package":
                self.fd.package = self.take()
                self.take(";")
            elif tok == "import":
                self.fd.dependency.append(self.take().strip('"'))
                self.take(";")
            elif tok == "message":
                self.message(self.fd.message_type.add())
            elif tok == "enum":
                self.enum(self.fd.enum_type.add())
            else:
                raise SyntaxError(tok)
        return self.fd

    def enum(self, e):
        e.name = self.take()
        self.take("{")
        while self.peek() != "}":
            v = e.value.add()
            v.name = self.take()
            self.take("=")
            v.number = int(self.take())
            self.take(";")
        self.take("}")

    def field(self, m, first, oneof_index=None):
        f = m.field.add()
        label = F.LABEL_OPTIONAL
        typ = first
        if first == "repeated":
            label = F.LABEL_REPEATED
            typ = self.take()
        if typ == "map":
            self.take("<")
            kt = self.take()
            self.take(",")
            vt = self.take()
            self.take(">")
            f.name = self.take()
            entry = m.nested_type.add()
            entry.name = "".join(w.capitalize() if i else w[:1].upper() + w[1:]
                                 for i, w in enumerate(f.name.split("_"))) + "Entry"
            entry.options.map_entry = True
            k = entry.field.add(name="key", number=1, label=F.LABEL_OPTIONAL)
            self.settype(k, kt)
            v = entry.field.add(name="value", number=2, label=F.LABEL_OPTIONAL)
            self.settype(v, vt)
            label = F.LABEL_REPEATED
            f.type = F.TYPE_MESSAGE
            f.type_name = entry.name  # resolved relative to the enclosing message
            self._map_entries.append((f, entry))
        else:
            f.name = self.take()
            self.settype(f, typ)
        f.label = label
        self.take("=")
        f.number = int(self.take())
        self.take(";")
        if oneof_index is not None:
            f.oneof_index = oneof_index

    def settype(self, f, typ):
        if typ in SCALARS:
            f.type = SCALARS[typ]
        else:
            f.type_name = typ  # message or enum: fixed up in resolve()
            f.type = F.TYPE_MESSAGE

    def message(self, m):
        m.name = self.take()
        self.take("{")
        self._map_entries = getattr(self, "_map_entries", [])
        while self.peek() != "}":
            tok = self.take()
            if tok == "message":
                self.message(m.nested_type.add())
            elif tok == "enum":
                self.enum(m.enum_type.add())
            elif tok == "oneof":
                idx = len(m.oneof_decl)
                m.oneof_decl.add(name=self.take())
                self.take("{")
                while self.peek() != "}":
                    self.field(m, self.take(), idx)
                self.take("}")
            else:
                self.field(m, tok)
        self.take("}")


def resolve(fds):
    """Turns the type names the .proto files use into fully qualified ones and decides message/enum."""
    msgs, enums = set(), set()

    def walk(prefix, m):
        full = prefix + "." + m.name
        msgs.add(full)
        for e in m.enum_type:
            enums.add(full + "." + e.name)
        for n in m.nested_type:
            walk(full, n)

    for fd in fds:
        for e in fd.enum_type:
            enums.add("." + fd.package + "." + e.name)
        for m in fd.message_type:
            walk("." + fd.package, m)
    msgs.add(".google.protobuf.Int32Value")

    def fix(scope, m):
        full = scope + "." + m.name
        for f in m.field:
            if not f.type_name:
                continue
            name = f.type_name
            cands = []
            parts = full.split(".")
            for cut in range(len(parts), 0, -1):
                cands.append(".".join(parts[:cut]) + "." + name)
            cands.append("." + name)
            for c in cands:
                if c in msgs:
                    f.type_name, f.type = c, F.TYPE_MESSAGE
                    break
                if c in enums:
                    f.type_name, f.type = c, F.TYPE_ENUM
                    break
            else:
                raise NameError("unresolved type %s in %s" % (name, full))
        for n in m.nested_type:
            fix(full, n)

    for fd in fds:
        for m in fd.message_type:
            fix("." + fd.package, m)


def build_pool():
    fds = [ProtoReader(n, p).parse() for n, p in FILES]
    resolve(fds)
    pool = descriptor_pool.Default()
    for fd in fds:
        pool.Add(fd)
    return pool, fds


def schema_lines(fds):
    """Same line format as dawn_b200_iir_schema()."""
    names = {1: "double", 2: "float", 5: "int32", 8: "bool", 9: "string"}
    out = []

    def tname(f):
        if f.type in names:
            return names[f.type]
        short = f.type_name.split(".")
        short = ".".join(short[4:]) if f.type_name.startswith(".dawn.proto") else short[-1]
        return ("enum " + short) if f.type == F.TYPE_ENUM else short

    def walk(m, prefix=""):
        entries = {n.name: n for n in m.nested_type if n.options.map_entry}
        for f in m.field:
            leaf = f.type_name.split(".")[-1] if f.type_name else ""
            if leaf in entries and f.type == F.TYPE_MESSAGE and f.type_name.endswith(m.name + "." + leaf):
                e = entries[leaf]
                t = "map<%s,%s>" % (tname(e.field[0]), tname(e.field[1]))
            else:
                t = ("repeated " if f.label == F.LABEL_REPEATED else "") + tname(f)
            out.append("%s %d %s %s" % (m.name, f.number, f.name, t))
        for e in m.enum_type:
            for v in e.value:
                out.append("enum %s.%s %s %d" % (m.name, e.name, v.name, v.number))
        for n in m.nested_type:
            if not n.options.map_entry:
                walk(n)

    for fd in fds:
        for e in fd.enum_type:
            for v in e.value:
                out.append("enum %s %s %d" % (e.name, v.name, v.number))
        for m in fd.message_type:
            walk(m)
    return sorted(out)


def main():
    pool, fds = build_pool()
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("dawn.proto.iir.StencilInstantiation"))
    outdir = os.path.join(ROOT, "tests", "golden", "proto")
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "schema.txt"), "w") as f:
        f.write("\n".join(schema_lines(fds)) + "\n")
    srcs = sorted(glob.glob(os.path.join(ROOT, "stencils", "*.iir")))
    srcs += sorted(glob.glob("/root/reference/dawn/test/unit-test/dawn/CodeGen/input/*.iir"))
    n = 0
    for path in srcs:
        msg = cls()
        try:
            json_format.Parse(open(path).read(), msg)
        except json_format.ParseError as e:  # fixtures written against an older schema revision
            print("skip %s: %s" % (os.path.basename(path), str(e)[:100]))
            continue
        name = os.path.splitext(os.path.basename(path))[0]
        with open(os.path.join(outdir, name + ".iirb"), "wb") as f:
            f.write(msg.SerializeToString(deterministic=True))
        # the official runtime's JSON rendering of the same message (original field names), for the
        # decode direction of the test
        with open(os.path.join(outdir, name + ".json"), "w") as f:
            f.write(json_format.MessageToJson(msg, preserving_proto_field_name=True, indent=None))
        n += 1
    print("wrote %d instantiations to %s" % (n, outdir))


if __name__ == "__main__":
    sys.exit(main())
