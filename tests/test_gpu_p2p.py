"""The path's only collective through the engine's own kernel: one-shot all-reduce(sum) over NVLink
peer memory (ctx_p2p_allreduce, csrc/ctx_p2p.cu) against NCCL -- correctness for both dtypes and
several message sizes, bit-identical results on every rank, both buffer slots / the epoch protocol
over 25 back-to-back calls, CUDA-graph replay.  Needs two GPUs on one node (skipped otherwise; the
single-GPU box of the driver's `-m gpu` run skips it, `scripts/runs/r2g.sh` / `r2h.sh` ran it at N = 2 / 8)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_peer_allreduce_matches_nccl_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(ROOT / "scripts" / "p2p_check.py")],
                       capture_output=True, text=True, timeout=300, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["ok"] and line["world"] == 2


def test_peer_allreduce_rejects_bad_arguments():
    import ctypes as C

    from catalax_b200 import engine

    L = engine.lib()
    h = C.c_void_p()
    buf = C.create_string_buffer(64)
    assert L.ctx_p2p_create(0, 17, 1024, C.byref(h), buf) == -1          # more than 16 ranks
    assert L.ctx_p2p_create(3, 2, 1024, C.byref(h), buf) == -1           # rank outside the world
    assert L.ctx_p2p_create(0, 1, 4096, C.byref(h), buf) == 0            # a world of one is legal
    assert L.ctx_p2p_allreduce(h, 0, 16, None, None) == -1               # not connected / null buffer
    assert L.ctx_p2p_connect(h, bytes(buf.raw)) == 0
    x = torch.arange(1000, dtype=torch.float32, device="cuda:0")
    assert L.ctx_p2p_allreduce(h, 0, x.numel(), C.c_void_p(x.data_ptr()),
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert L.ctx_p2p_allreduce(h, 0, 5000, C.c_void_p(x.data_ptr()), None) == -1      # larger than the buffer
    torch.cuda.synchronize()
    assert torch.equal(x, torch.arange(1000, dtype=torch.float32, device="cuda:0"))     # sum over one rank
    L.ctx_p2p_free(h)
