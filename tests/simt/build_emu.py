"""Build tests/simt/_build/libthb_lane_emu.so: theseus_b200/csrc/thb_sparse_lane.cu compiled for the HOST with g++ on top of
simt_shim.h (one OS thread per CUDA thread).  The source is used as it is except for three mechanical rewrites done here:
  kernel<<<grid, block, smem, stream>>>(args)  ->  SIMT_LAUNCH(kernel, grid, block, smem, args)
  extern __shared__ T name[];                   ->  T* name = simt_dyn_smem<T>();
  #include "thb_common.cuh"                     ->  #include "simt_shim.h"
Test infrastructure only (like oracle/): nothing in the product imports it."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "..", "theseus_b200", "csrc", "thb_sparse_lane.cu")
OUT_DIR = os.path.join(HERE, "_build")


def _split_top_level(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[" :
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite(text: str) -> str:
    out = []
    for line in text.split("\n"):
        m = re.search(r"([\w:]+(?:<[\w\s,:]*>)?)<<<(.*)>>>\((.*)\)", line)
        if m:
            kernel, cfg, args = m.group(1), _split_top_level(m.group(2)), m.group(3)
            assert len(cfg) == 4, line
            line = line[:m.start()] + f"SIMT_LAUNCH(({kernel}), {cfg[0]}, {cfg[1]}, {cfg[2]}, {args})" + line[m.end():]
        m = re.search(r"extern __shared__ (\w+) (\w+)\[\];", line)
        if m:
            line = line[:m.start()] + f"{m.group(1)}* {m.group(2)} = simt_dyn_smem<{m.group(1)}>();" + line[m.end():]
        line = line.replace('#include "thb_common.cuh"', '#include "simt_shim.h"')
        out.append(line)
    text = "\n".join(out)
    assert "<<<" not in text.replace("<<<...>>>", "") and "extern __shared__" not in text
    return text


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = os.path.join(OUT_DIR, "libthb_lane_emu.so")
    gen = os.path.join(OUT_DIR, "thb_sparse_lane_emu.cpp")
    src_text = rewrite(open(SRC).read())
    if not force and os.path.exists(lib) and os.path.exists(gen) and open(gen).read() == src_text \
            and os.path.getmtime(lib) >= os.path.getmtime(os.path.join(HERE, "simt_shim.h")):
        return lib
    open(gen, "w").write(src_text)
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-DTHB_SIMT_EMU", "-Wno-unknown-pragmas",
           # the CUDA library exports host stubs with the SAME mangled kernel names: bind this library's references to its own definitions
           "-Wl,-Bsymbolic", "-fno-semantic-interposition", "-I", HERE, gen, "-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on the emulation build:\n" + r.stdout + r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force=True))
