#!/usr/bin/env python3
"""Turns the reference's GLSL shaders into C++ translation units that g++ can compile against glsl_shim.h.
TEST INFRASTRUCTURE ONLY (oracle/).  Reads the shader sources where they lie (default /root/reference/src/shaders),
writes ONLY into the output directory (oracle/_ref/gen, git-ignored).  Nothing of the reference is committed.

Only declarations C++ cannot parse are rewritten; every statement inside a function body is left as written:
  #version / #extension                      dropped
  #include "x.h"                             x.h is rewritten the same way into the output directory
  #include "../config.h"                     absolute path of the reference's config.h (plain C macros, untouched)
  1.5 / .5 / 2.0                             1.5f / .5f / 2.0f       (GLSL literals are 32-bit floats)
  out T name  (parameters)                   T& name
  layout(constant_id = N) const bool X = v   static bool X = v        (+ rs_spec)
  layout(local_size_x = ..) in               static const uint rs_local[3]
  layout(push_constant) uniform block {..}   static struct rs_pc + a reference per member
  layout(binding = N) buffer B { T a[]; }    static RobustArray<T> a  (zero on out-of-range reads, writes discarded)
  layout(binding = N) buffer B { uint a; ..} pointer to the block, members reachable through macros
  layout(binding = N) uniform texture2D / sampler / image2D    static objects (+ rs_bind)
  shared T x / taskPayloadSharedEXT T x      static thread_local   (one workgroup runs on one host thread)
  layout(triangles, ...) out                 dropped;  layout(location = N) out T x[]  ->  static thread_local T x[256]
  layout(location = N) out T x  (vertex stage)   static thread_local T x
  void main()                                static void rs_main()
"""
import os
import re
import sys

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def common_rewrites(text, src_dir):
    text = re.sub(r"^[ \t]*#(version|extension)[^\n]*\n", "", text, flags=re.M)
    text = text.replace('#include "../config.h"', '#include "%s"' % os.path.normpath(os.path.join(src_dir, "..", "config.h")))
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = re.sub(r"([(,]\s*)out\s+(\w+)\s+(\w+)", r"\1\2& \3", text)  # only inside parameter lists
    return text


def members_of(body):
    """[(type, name, is_unsized_array)] of a block body."""
    out = []
    for decl in body.split(";"):
        decl = re.sub(r"//[^\n]*", "", decl).strip()
        if not decl:
            continue
        m = re.match(r"^(\w+)\s+(.+)$", decl, flags=re.S)
        typ, rest = m.group(1), m.group(2)
        for item in rest.split(","):
            item = item.strip()
            mm = re.match(r"^(\w+)\s*(\[\s*\])?$", item)
            if not mm:
                raise SystemExit("gen.py: cannot parse block member %r" % decl)
            out.append((typ, mm.group(1), mm.group(2) is not None))
    return out


def rewrite_shader(text, name):
    binds, specs, defines = [], [], []
    state = {"pc": False, "payload": None, "barrier": "barrier()" in text}

    def spec(m):
        specs.append((int(m.group(1)), m.group(2)))
        return "static bool %s = %s;" % (m.group(2), m.group(3))

    text = re.sub(r"layout\s*\(\s*constant_id\s*=\s*(\d+)\s*\)\s*const\s+bool\s+(\w+)\s*=\s*(\w+)\s*;", spec, text)

    def local(m):
        return "static const uint rs_local[3] = { %s, %s, %s };" % (m.group(1), m.group(2), m.group(3))

    text, n = re.subn(r"layout\s*\(\s*local_size_x\s*=\s*([^,]+?)\s*,\s*local_size_y\s*=\s*([^,]+?)\s*,\s*local_size_z\s*=\s*([^)]+?)\s*\)\s*in\s*;", local, text)
    assert n <= 1, "local size of %s" % name
    has_local = n == 1

    def push(m):
        state["pc"] = True
        mem = members_of(m.group(2))
        refs = "".join("\nstatic %s& %s = rs_pc.%s;" % (typ, member, member) for typ, member, _ in mem)
        return "static struct rs_pc_t {%s} rs_pc;%s" % (m.group(2), refs)

    text = re.sub(r"layout\s*\(\s*push_constant\s*\)\s*uniform\s+(\w+)\s*\{(.*?)\}\s*;", push, text, flags=re.S)

    def buffer(m):
        binding, block, body = int(m.group(1)), m.group(3), m.group(4)
        mem = members_of(body)
        if all(unsized for _, _, unsized in mem):
            assert len(mem) == 1
            typ, member, _ = mem[0]
            binds.append((binding, "%s.bind(p, bytes);" % member))
            return "static RobustArray<%s> %s; // block %s" % (typ, member, block)
        assert not any(unsized for _, _, unsized in mem), "mixed block %s" % block
        for _, member, _ in mem:
            defines.append("#define %s rs_b_%s->%s" % (member, block, member))
        binds.append((binding, "rs_b_%s = static_cast<%s_block*>(p);" % (block, block)))
        return "struct %s_block {%s};\nstatic %s_block* rs_b_%s;" % (block, body, block, block)

    text = re.sub(r"layout\s*\(\s*binding\s*=\s*(\d+)\s*\)\s*(readonly\s+|writeonly\s+)?buffer\s+(\w+)\s*\{(.*?)\}\s*;", buffer, text, flags=re.S)

    def resource(m):
        binding, typ, var = int(m.group(1)), m.group(3), m.group(4)
        if typ == "sampler":
            binds.append((binding, "(void)p;"))
        else:
            binds.append((binding, "%s = *static_cast<const %s*>(p);" % (var, typ)))
        return "static %s %s;" % (typ, var)

    text = re.sub(r"layout\s*\(\s*binding\s*=\s*(\d+)\s*(?:,\s*\w+\s*)?\)\s*uniform\s+(writeonly\s+|readonly\s+)?(texture2D|sampler|image2D)\s+(\w+)\s*;", resource, text)

    # mesh stage: the output topology declaration carries no code; per-vertex outputs become plain arrays
    text = re.sub(r"layout\s*\(\s*triangles\s*,[^)]*\)\s*out\s*;", "", text)
    outputs = []

    def output(m):
        outputs.append((int(m.group(1)), m.group(3)))
        return "static thread_local %s %s[256];" % (m.group(2), m.group(3))

    text = re.sub(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)\s*out\s+(?:flat\s+)?(\w+)\s+(\w+)\s*\[\s*\]\s*;", output, text)

    def output_scalar(m):  # vertex stage: one value per invocation
        outputs.append((int(m.group(1)), "&" + m.group(3)))
        return "static thread_local %s %s;" % (m.group(2), m.group(3))

    text = re.sub(r"layout\s*\(\s*location\s*=\s*(\d+)\s*\)\s*out\s+(?:flat\s+)?(\w+)\s+(\w+)\s*;", output_scalar, text)

    def payload(m):
        state["payload"] = m.group(2)
        return "static thread_local %s %s;" % (m.group(1), m.group(2))

    text = re.sub(r"\btaskPayloadSharedEXT\s+(\w+)\s+(\w+)\s*;", payload, text)
    text = re.sub(r"^shared\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text, flags=re.M)
    text, n = re.subn(r"\bvoid\s+main\s*\(\s*\)", "static void rs_main()", text)
    assert n == 1, "main of %s" % name
    if "layout" in re.sub(r"//[^\n]*", "", text):
        raise SystemExit("gen.py: unhandled layout declaration left in %s" % name)

    tail = ["", "" if has_local else "static const uint rs_local[3] = { 1, 1, 1 }; // not a compute-like stage: one invocation per dispatch element",
            "static void rs_bind(int binding, void* p, size_t bytes)", "{", "\t(void)bytes;", "\tswitch (binding)", "\t{"]
    for b in sorted(set(x[0] for x in binds)):
        tail.append("\tcase %d:" % b)
        for bb, stmt in binds:
            if bb == b:
                tail.append("\t\t" + stmt)
        tail.append("\t\tbreak;")
    tail += ["\tdefault:", "\t\tbreak;", "\t}", "}", ""]
    tail += ["static void rs_push(const void* p, size_t bytes)", "{"]
    tail += ["\tmemcpy(&rs_pc, p, bytes < sizeof(rs_pc) ? bytes : sizeof(rs_pc));"] if state["pc"] else ["\t(void)p; (void)bytes;"]
    tail += ["}", "", "static void rs_spec(int id, int value)", "{", "\t(void)id; (void)value;"]
    for sid, var in specs:
        tail.append("\tif (id == %d) %s = value != 0;" % (sid, var))
    tail += ["}", ""]
    tail.append("static void* rs_payload() { return %s; }" % ("&" + state["payload"] if state["payload"] else "nullptr"))
    tail += ["", "static void* rs_output(int location)", "{", "\t(void)location;"]
    for loc, var in outputs:
        tail.append("\tif (location == %d) return %s;" % (loc, var))
    tail += ["\treturn nullptr;", "}"]
    # the member macros must touch neither the declarations above them nor the glue: both go right before rs_main
    text = text.replace("static void rs_main()", "\n".join(tail) + "\n\n" + "\n".join(defines) + "\n\nstatic void rs_main()", 1)
    return text, state


def main():
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    for header in ("mesh.h", "math.h"):
        with open(os.path.join(src_dir, header)) as f:
            text = common_rewrites(f.read(), src_dir)
        with open(os.path.join(out_dir, header), "w") as f:
            f.write("// generated by oracle/refshader/gen.py from %s -- build artefact, never committed\n" % os.path.join(src_dir, header) + text)
    for shader in sys.argv[3:]:
        name = shader.replace(".comp.glsl", "").replace(".glsl", "").replace(".", "_")
        with open(os.path.join(src_dir, shader)) as f:
            text = common_rewrites(f.read(), src_dir)
        text, state = rewrite_shader(text, name)
        with open(os.path.join(out_dir, name + ".cpp"), "w") as f:
            f.write("// generated by oracle/refshader/gen.py from %s -- build artefact, never committed\n" % os.path.join(src_dir, shader))
            f.write('#include "glsl_shim.h"\n#include "rs_shader.h"\n\nnamespace glsl\n{\nnamespace rs_%s\n{\n\n' % name)
            f.write(text)
            f.write("\n} // namespace rs_%s\n} // namespace glsl\n\n" % name)
            f.write('extern const RsShader rs_shader_%s = { "%s", glsl::rs_%s::rs_local, glsl::rs_%s::rs_main, glsl::rs_%s::rs_bind, glsl::rs_%s::rs_push, glsl::rs_%s::rs_spec, glsl::rs_%s::rs_payload, glsl::rs_%s::rs_output, %d };\n' % (name, shader, name, name, name, name, name, name, name, 1 if state["barrier"] else 0))


if __name__ == "__main__":
    main()
