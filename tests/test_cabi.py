"""The C-ABI library loads and exports every symbol include/ctx_b200.h declares (no GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "ctx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for s in ("ctx_model_compile", "ctx_model_free", "ctx_solve", "ctx_solve_sens",
              "ctx_solve_host", "ctx_workspace_bytes", "ctx_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(ROOT / "catalax_b200" / "libctxb200.so"))
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.ctx_abi_version.restype = ctypes.c_int
    assert lib.ctx_abi_version() == 4


def test_no_cpu_fallback_without_gpu():
    """Compute entry points must fail loudly (CTX_E_NOGPU) when no device is visible."""
    import numpy as np

    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    if engine.device_count() > 0:
        pytest.skip("a GPU is visible")
    field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
    m = engine.CompiledModel(field, "float64")
    y0, th, c = cases.mm_inputs(4, dtype=np.float64)
    with pytest.raises(engine.EngineError, match="no CPU fallback|no CUDA"):
        m.solve_host(y0, th, c, np.linspace(0, 1, 3), in_axes=(0, 0, 0, None))


def test_descriptor_mismatch_is_rejected():
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    field = cg.generate_field(cases.field_spec(cases.michaelis_menten()))
    field.S = 4  # lie about the number of states
    with pytest.raises(engine.EngineError, match="descriptor does not match"):
        engine.CompiledModel(field, "float32")


def test_nvrtc_compiles_all_variants_for_sm100a_without_gpu():
    from catalax_b200 import engine
    from catalax_b200.tools import codegen as cg
    from tests import cases

    field = cg.generate_field(cases.field_spec(cases.michaelis_menten()), sens_theta=True)
    m = engine.CompiledModel(field, "float32")
    for solver in ("tsit5", "dopri5", "euler", "heun"):
        cubin = m.cubin(solver, tangents=(solver == "tsit5"))
        assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000


def test_mlp_kernel_sass_uses_tcgen05_and_tmem():
    """Evidence that K3 is Blackwell-native: UTC*MMA (tcgen05.mma) and LDTM (tcgen05.ld) in the
    SASS of the NVRTC-compiled kernel (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import subprocess
    import tempfile

    from catalax_b200 import engine
    from catalax_b200.neural import NeuralODE
    from catalax_b200.tools import codegen_mlp as cm

    net = NeuralODE(4, 64, 3, ["a", "b", "c", "d"], activation="tanh", max_time=10.0, key=3)
    m = engine.CompiledModel(cm.generate_mlp_field(net.spec), "float32")
    cubin = m.cubin("tsit5", False)
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cubin); f.flush()
        sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    assert "ctx_solve_mlp_tc" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTCBAR" in sass
    assert "HMMA.16" not in sass  # no legacy mma.sync path
