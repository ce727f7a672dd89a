"""In-tree build of libctxb200.so (nvcc, sm_100a).  Used by __graft_entry__.build()."""
from __future__ import annotations

import hashlib
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libctxb200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _embed(header: Path, out: Path, symbol: str) -> None:
    """Turn the integrator template into a C string the library hands to NVRTC."""
    data = header.read_bytes()
    lines = [f"// generated from {header.name} by catalax_b200/_build.py -- do not edit",
             f"extern const char {symbol}[];", f"const char {symbol}[] = {{"]
    for i in range(0, len(data), 20):
        lines.append("  " + ",".join(str(b) for b in data[i:i + 20]) + ",")
    lines.append("  0};")
    out.write_text("\n".join(lines) + "\n")


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "ctx_b200.h"]


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


_MARK = b"CTXB200-SOURCE-STAMP:"


def lib_stamp(path: Path = LIB):
    """Source hash baked into a built library (`ctx_build_stamp()`), read from the file itself:
    the library, not a side file that git can update behind its back, says what it was built
    from.  None when the file is missing or carries no stamp."""
    try:
        data = path.read_bytes()
    except OSError:
        return None
    i = data.find(_MARK)
    if i < 0:
        return None
    return data[i + len(_MARK): i + len(_MARK) + 64].decode("ascii", "replace")


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = _stamp()
    if not force and lib_stamp() == stamp:
        return LIB
    gen = CSRC / "_gen"
    gen.mkdir(exist_ok=True)
    _embed(CSRC / "ctx_integrator.cuh", gen / "ctx_integrator_embed.cpp", "ctx_integrator_src")
    _embed(CSRC / "ctx_mlp_tc.cuh", gen / "ctx_mlp_tc_embed.cpp", "ctx_mlp_tc_src")
    _embed(CSRC / "ctx_mlp_adjoint.cuh", gen / "ctx_mlp_adjoint_embed.cpp", "ctx_mlp_adjoint_src")
    _embed(CSRC / "ctx_solve_v2.cuh", gen / "ctx_solve_v2_embed.cpp", "ctx_solve_v2_src")
    cu = [str(p) for p in sorted(CSRC.glob("*.cu"))]
    cmd = [NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
           f'-DCTX_BUILD_STAMP="{stamp}"',
           *cu, str(gen / "ctx_integrator_embed.cpp"), str(gen / "ctx_mlp_tc_embed.cpp"),
           str(gen / "ctx_mlp_adjoint_embed.cpp"), str(gen / "ctx_solve_v2_embed.cpp"), "-o", str(LIB), "-ldl",
           "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    tmp = LIB.with_suffix(".so.tmp")
    cmd[cmd.index(str(LIB))] = str(tmp)
    subprocess.run(cmd, check=True, cwd=str(CSRC))
    os.replace(tmp, LIB)
    assert lib_stamp() == stamp, "built library does not carry the source stamp"
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
