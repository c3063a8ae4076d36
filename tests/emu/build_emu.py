"""Builds tests/emu/libfestim_b200_emu.so: the library's own sources (festim_b200/csrc/*.cu)
compiled with g++ against the CPU emulation of the CUDA execution model (cuda_emu.h), with
every ``kernel<<<grid, block, smem, stream>>>(args)`` rewritten to ``EMU_KERNEL((kernel), grid,
block, smem, stream)(args)``.  The C ABI is the product's, so the CPU tests drive the real
kernel code (assembly, SpMV, vector kernels) through the same entry points.  Test
infrastructure only."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "festim_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libfestim_b200_emu.so")
SOURCES = ["api.cu", "assemble.cu", "linalg.cu", "newton.cu"]

LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
EXTERN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def transform(text):
    text = LAUNCH.sub(lambda m: f"EMU_KERNEL(({m.group(1)}), {m.group(2)})(", text)
    text = EXTERN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)emu_dynamic_smem();", text)
    # FB_LAUNCH(ctx, kernel, grid, block, smem, args...) expands to a <<<>>> launch
    text = text.replace("kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);",
                        "EMU_KERNEL((kernel), (grid), (block), (smem))(__VA_ARGS__);")
    # inline PTX has no host meaning: the emulation gives plain loads / stores / atomics
    text = re.sub(r'asm volatile\("ld\.[^"]*"\s*:\s*"=[rl]"\((\w+)\)\s*:\s*"l"\((\w+)\)\s*:\s*"memory"\);',
                  r"\1 = *\2;", text)
    text = re.sub(r'asm volatile\("st\.[^"]*"\s*::\s*"l"\(([^)]*(?:\([^)]*\))?[^)]*)\),\s*"[rl]"\((.*?)\)\s*:\s*"memory"\);',
                  r"*(\1) = (\2);", text)
    text = re.sub(r'asm volatile\("atom\.add[^"]*"\s*:\s*"=r"\((\w+)\)\s*:\s*"l"\((\w+)\),\s*"r"\((\w+)\)\s*:\s*"memory"\);',
                  r"\1 = atomicAdd(\2, \3);", text)
    return text


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "build_emu.py")] + \
           [os.path.join(ROOT, "include", "festim_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not (force or stale()):
        return LIB
    os.makedirs(os.path.join(OUT, "csrc"), exist_ok=True)
    cpps = []
    for name in os.listdir(CSRC):
        src = open(os.path.join(CSRC, name)).read()
        out = transform(src)
        if name.endswith(".cu"):
            out = '#include "cuda_emu.h"\n' + out
            dst = os.path.join(OUT, "csrc", name[:-3] + ".cpp")
            if name in SOURCES:
                cpps.append(dst)
        else:
            dst = os.path.join(OUT, "csrc", name)
        with open(dst, "w") as fh:
            fh.write(out)
    # the sources include "../../include/festim_b200.h" relative to csrc/
    os.makedirs(os.path.join(HERE, "include"), exist_ok=True)
    hdr = os.path.join(HERE, "include", "festim_b200.h")
    with open(hdr, "w") as fh:
        fh.write(open(os.path.join(ROOT, "include", "festim_b200.h")).read())
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-DFB_EMU", "-Wl,-Bsymbolic", "-I", HERE, "-I", cuda_inc,
           "-o", LIB, os.path.join(HERE, "cuda_emu.cpp")] + cpps + ["-ldl"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
