#!/usr/bin/env python
"""Compile the REFERENCE'S OWN compute shaders for the CPU: oracle/_ref/libhg_refshaders.so.

TEST INFRASTRUCTURE (like everything under oracle/): the library is what the CPU restatement
(oracle/hg_oracle.c) is validated against, dispatch by dispatch and bit for bit
(tests/test_refshaders.py).  Nothing in the product loads it.

The shader sources are read where they lie (/root/reference/glsl/*.glsl, or $HG_REFERENCE_GLSL); none
of their text is stored in this repository.  Each shader is turned into a C++ namespace by purely
lexical steps and compiled by g++ -std=c++20 -ffp-contract=off against oracle/refshader/glsl_shim.hpp,
which supplies the GLSL vocabulary (vector types with swizzles, built-ins, texelFetch/imageLoad/
imageStore on RGBA32F arrays, gl_GlobalInvocationID ...).  The lexical steps:

  * `#include <name>` is replaced by that file (simplex_noise: only the functions the shader reaches),
    `#version`, `#line` and `layout(local_size...) in;` are dropped;
  * `layout(...) uniform sampler2D|image2D NAME;` becomes a bindable image variable NAME,
    `layout(std140, ...) uniform BLOCK { TYPE NAME; };` becomes a variable NAME of TYPE,
    `uniform int|float NAME;` a plain variable;
  * floating literals get an `f` suffix (a GLSL literal is a float, a C++ one a double);
  * `main` is renamed; the generated dispatcher runs it once per invocation of the grid.

Only the .so is written to oracle/_ref/ (git-ignored, but it travels to the GPU box with the other
built files); the generated C++ lives in a temporary directory."""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
GLSL = os.environ.get("HG_REFERENCE_GLSL", "/root/reference/glsl")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(OUT_DIR, "libhg_refshaders.so")

# shader -> functions / structs / defines of simplex_noise.glsl it reaches (that file is a library of
# 40 functions; the rest would only enlarge the vocabulary the shim has to provide)
SHADERS = {
    "hydro_flux": [], "hydro_erosion": [], "sediment_transport": [], "thermal_erosion": [],
    "thermal_transport": [], "smoothing": [],
    "rain": ["MAX_FBM_ITERATIONS", "gln_tFBMOpts", "gln_rand3", "gln_simplex", "gln_sfbm"],
    "particle": [], "particle_erosion": [],
    "heightmap": ["MAX_FBM_ITERATIONS", "gln_tFBMOpts", "gln_rand3", "gln_simplex", "gln_sfbm", "hash", "noised", "perlfbm", "erosion_perlfbm"],
}

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(f?)(?![\w.])")


def read(name):
    with open(os.path.join(GLSL, name + ".glsl")) as f:
        return f.read()


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def slice_items(src, names):
    """definitions of the named #defines / structs / functions from a GLSL library file, in file order"""
    src = strip_comments(src)
    out = []
    for name in names:
        m = re.search(r"^[ \t]*#define[ \t]+%s\b[^\n]*" % re.escape(name), src, flags=re.M)
        if m:
            out.append((m.start(), m.group(0)))
            continue
        m = re.search(r"^[ \t]*(struct[ \t]+%s\b|[A-Za-z_]\w*[ \t]+%s[ \t]*\()" % (re.escape(name), re.escape(name)), src, flags=re.M)
        if not m:
            raise SystemExit(f"{name} not found in the noise library")
        i = src.index("{", m.start())
        depth, j = 0, i
        while True:
            depth += src[j] == "{"
            depth -= src[j] == "}"
            j += 1
            if depth == 0:
                break
        text = src[m.start():j]
        if text.lstrip().startswith("struct"):
            text += ";"
        out.append((m.start(), text))
    return "\n".join(t for _, t in sorted(out))


def translate(shader):
    src = strip_comments(read(shader))
    binds, blocks, plain, buffers = [], [], [], []

    def include(m):
        name = m.group(1)
        if name == "simplex_noise":
            return slice_items(read(name), SHADERS[shader])
        return strip_comments(read(name))
    src = re.sub(r"^[ \t]*#include[ \t]*<(\w+)>[^\n]*", include, src, flags=re.M)
    src = re.sub(r"^[ \t]*#(version|line)[^\n]*", "", src, flags=re.M)
    src = re.sub(r"layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;", "", src)

    def image(m):
        binds.append(m.group(2))
        return f"{m.group(1)} {m.group(2)};"
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(?:(?:readonly|writeonly|volatile|coherent|restrict)\s+)*(sampler2D|image2D|uimage2D)\s+(\w+)\s*;", image, src)

    def block(m):
        blocks.append((m.group(1), m.group(2)))
        return f"{m.group(1)} {m.group(2)};"
    src = re.sub(r"layout\s*\(\s*std140[^)]*\)\s*uniform\s+\w+\s*\{\s*(\w+)\s+(\w+)\s*;\s*\}\s*;", block, src)

    # the droplet SSBO (heightmap.glsl only clears it under #if defined(PARTICLE_COUNT), which the grid build does not define)
    def ssbo(m):
        buffers.append((m.group(1), m.group(2), m.group(3)))
        return f"{m.group(2)}* {m.group(3)} = nullptr;"
    src = re.sub(r"layout\s*\(\s*std430[^)]*\)\s*buffer\s+(\w+)\s*\{\s*(\w+)\s+(\w+)\s*\[\s*\]\s*;\s*\}\s*;", ssbo, src)

    def uniform(m):
        plain.append((m.group(1), m.group(2)))
        return f"{m.group(1)} {m.group(2)};"
    src = re.sub(r"^[ \t]*uniform\s+(int|float|bool|uint)\s+(\w+)\s*;", uniform, src, flags=re.M)
    if re.search(r"\blayout\s*\(|\buniform\b|\bbuffer\b", src):
        raise SystemExit(f"{shader}: a declaration the translator does not know:\n" + "\n".join(l for l in src.split("\n") if re.search(r"layout|uniform|buffer", l)))
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    return src, binds, blocks, plain, buffers


def generate():
    parts = ['#define GL_core_profile 1\n#include "glsl_shim.hpp"\n#include <cstring>\n',
             "namespace glsl { thread_local U3 gl_GlobalInvocationID, gl_NumWorkGroups, gl_WorkGroupSize; }\n"]
    table = []
    for k, shader in enumerate(SHADERS):
        src, binds, blocks, plain, buffers = translate(shader)
        # bindings.glsl has an include guard: lift it so that every namespace gets its own copy
        src = src.replace("HYDR_GL_BINDINGS_HPP", f"HYDR_GL_BINDINGS_{k}")
        parts.append(f"namespace glsl {{ namespace ref_{shader} {{\n#define bool gbool\n{src}\n#undef bool\n")
        parts.append("static int bind(const char* n, float* p, int w, int h) {\n")
        for b in binds:
            parts.append(f'    if (!strcmp(n, "{b}")) {{ {b}.p = (decltype({b}.p))p; {b}.w = w; {b}.h = h; return 0; }}\n')
        for blk, t, n in buffers:
            parts.append(f'    if (!strcmp(n, "{blk}")) {{ {n} = ({t}*)p; return 0; }}\n')
        parts.append("    return -1;\n}\nstatic int set_uniform(const char* n, const void* p, int bytes) {\n")
        for t, n in blocks + plain:
            parts.append(f'    if (!strcmp(n, "{n}")) {{ if (bytes != (int)sizeof({n})) return -2; memcpy(&{n}, p, sizeof({n})); return 0; }}\n')
        parts.append("    return -1;\n}\n")
        parts.append("static int run1d(int count) {\n"
                     "    gl_WorkGroupSize.x = WRKGRP_SIZE_X * WRKGRP_SIZE_Y; gl_WorkGroupSize.y = 1; gl_WorkGroupSize.z = 1;\n"
                     "    gl_NumWorkGroups.x = count / (WRKGRP_SIZE_X * WRKGRP_SIZE_Y); gl_NumWorkGroups.y = 1; gl_NumWorkGroups.z = 1;\n"
                     "    for (int id = 0; id < count; id++) {\n"
                     "        gl_GlobalInvocationID.x = id; gl_GlobalInvocationID.y = 0; gl_GlobalInvocationID.z = 0;\n"
                     "        gl_GlobalInvocationID.xy = uvec2(id, 0);\n"
                     "        shader_main();\n    }\n    return 0;\n}\n")
        # grid dispatch: invocations are independent (each writes only its own texel of the write images), so rows run
        # on all host threads; the built-in variables are thread_local
        parts.append("static int run(int W, int H) {\n"
                     "    _Pragma(\"omp parallel for schedule(static)\")\n"
                     "    for (int y = 0; y < H; y++) {\n"
                     "        gl_WorkGroupSize.x = WRKGRP_SIZE_X; gl_WorkGroupSize.y = WRKGRP_SIZE_Y; gl_WorkGroupSize.z = 1;\n"
                     "        gl_NumWorkGroups.x = W / WRKGRP_SIZE_X; gl_NumWorkGroups.y = H / WRKGRP_SIZE_Y; gl_NumWorkGroups.z = 1;\n"
                     "        for (int x = 0; x < W; x++) {\n"
                     "            gl_GlobalInvocationID.x = x; gl_GlobalInvocationID.y = y; gl_GlobalInvocationID.z = 0;\n"
                     "            gl_GlobalInvocationID.xy = uvec2(x, y);\n"
                     "            shader_main();\n        }\n    }\n    return 0;\n}\n")
        parts.append("static_assert(sizeof(Erosion_data) == 96 && sizeof(Rain_data) == 20 && sizeof(Map_settings_data) == 96 && sizeof(Particle) == 48, \"std140 / std430 images of the blocks\");\n")
        parts.append("} }\n")
        table.append(shader)
    parts.append('extern "C" {\n')
    for fn, sig, call in (("ref_bind", "const char* s, const char* n, float* p, int w, int h", "bind(n, p, w, h)"),
                          ("ref_set_uniform", "const char* s, const char* n, const void* p, int bytes", "set_uniform(n, p, bytes)"),
                          ("ref_run", "const char* s, int W, int H", "run(W, H)"),
                          ("ref_run1d", "const char* s, int count", "run1d(count)")):
        parts.append(f"int {fn}({sig}) {{\n")
        for shader in table:
            parts.append(f'    if (!strcmp(s, "{shader}")) return glsl::ref_{shader}::{call};\n')
        parts.append("    return -3;\n}\n")
    parts.append("}\n")
    return "".join(parts)


def build():
    if not os.path.isdir(GLSL):
        raise SystemExit(f"{GLSL} not found: the reference shaders are needed to build {OUT}")
    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "ref_shaders.cpp")
        with open(cpp, "w") as f:
            f.write(generate())
        if "--keep" in sys.argv:
            import shutil
            shutil.copy(cpp, "/tmp/ref_shaders.cpp")
        cmd = ["g++", "-std=c++20", "-O3", "-march=x86-64-v3", "-ffp-contract=off", "-fno-strict-aliasing", "-fopenmp", "-ftls-model=initial-exec", "-fPIC", "-shared",
               "-Wno-narrowing", "-I", HERE, "-o", OUT, cpp, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr[:6000])
            raise SystemExit("g++ failed on the translated shaders")
    print("built", OUT)


if __name__ == "__main__":
    build()
