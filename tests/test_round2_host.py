"""Host-side logic added in round 2 (no GPU): the XLA custom-call descriptors, the source stamp
inside the library, checkpoint round trips of all three neural classes (+ RBF), the model JSON
writer, the sampler's constraining transforms and the posterior helper functions."""
import ctypes as C
import json
import math

import numpy as np
import pytest

import catalax_b200 as ctx
from catalax_b200 import _build, engine

GOLD = __import__("pathlib").Path(__file__).parent / "golden" / "reference_fixtures"


def test_xla_descriptors_match_the_library_layout():
    L = engine.lib()
    assert L.ctx_abi_version() == engine.ABI_VERSION == 4
    assert C.sizeof(engine.ctx_xla_solve_desc) == L.ctx_xla_desc_bytes(0)
    assert C.sizeof(engine.ctx_xla_loglik_desc) == L.ctx_xla_desc_bytes(1)
    assert L.ctx_xla_desc_bytes(7) == 0
    # a malformed opaque is parked as an error, not a crash (no GPU is touched)
    assert L.ctx_xla_last_status() == 0
    arr = (C.c_void_p * 9)()
    L.ctx_xla_solve(None, arr, b"xx", 2)
    assert L.ctx_xla_last_status() == -1            # CTX_E_INVALID
    assert L.ctx_xla_last_status() == 0             # reading clears the slot
    buf = C.create_string_buffer(256)
    L.ctx_xla_last_error(buf, 256)
    assert b"opaque descriptor" in buf.value
    d = engine.ctx_xla_solve_desc()
    d.abi, d.model = 2, 1                           # stale ABI version in the descriptor
    L.ctx_xla_solve(None, arr, bytes(d), C.sizeof(d))
    assert L.ctx_xla_last_status() == -1


def test_library_carries_the_stamp_of_its_sources(tmp_path):
    assert _build.lib_stamp() == _build._stamp()
    assert engine.lib().ctx_build_stamp().decode() == _build._stamp()
    # the stamp is read from the file itself: a library without one (or a missing file) never matches
    fake = tmp_path / "lib.so"
    fake.write_bytes(b"\x7fELF" + b"\0" * 64)
    assert _build.lib_stamp(fake) is None and _build.lib_stamp(tmp_path / "absent.so") is None
    fake.write_bytes(b"junk" + _build._MARK + b"0" * 64 + b"tail")
    assert _build.lib_stamp(fake) == "0" * 64 != _build._stamp()
    # neither generated artefact is tracked any more (a tracked stamp defeated the check)
    import subprocess
    tracked = subprocess.run(["git", "ls-files", "catalax_b200/csrc"], capture_output=True, text=True,
                             cwd=str(_build.HERE.parent)).stdout
    assert ".build_stamp" not in tracked and "_gen/" not in tracked


def test_neuralode_defaults_follow_the_reference():
    from catalax_b200.neural import NeuralODE, RateFlowODE

    # neuralbase.py:79: NeuralODE's default activation is tanh (softplus is only the inner MLP's)
    assert NeuralODE(2, 4, 1, ["a", "b"]).spec.activation == "tanh"
    assert RateFlowODE(2, 2, 4, 1, ["a", "b"]).spec.activation == "celu"


def test_checkpoint_without_activation_falls_back_to_the_class_default(tmp_path):
    from catalax_b200.neural import NeuralODE

    net = NeuralODE(2, 4, 1, ["a", "b"], activation="celu", max_time=3.0, key=1)
    path = net.save_to_eqx(tmp_path, "m")
    raw = open(path, "rb").read()
    head, rest = raw.split(b"\n", 1)
    hp = json.loads(head)["hyperparameters"]
    assert hp["activation"] == "celu"
    del hp["activation"]
    open(path, "wb").write(json.dumps({"hyperparameters": hp}).encode() + b"\n" + rest)
    assert NeuralODE.from_eqx(path).spec.activation == "tanh"      # neuralbase.py:514-518


def test_eqx_round_trips_of_every_class(tmp_path):
    from catalax_b200.neural import NeuralODE, RateFlowODE, RBFLayer, UniversalODE

    f32 = lambda a: np.asarray(a, dtype=np.float32)
    rf = RateFlowODE(4, 3, 8, 2, ["a", "b", "c", "d"], key=1, max_time=7.5,
                     mass_constraint=np.array([[1.0, 1.0, 0.0, 0.0]]))
    rf2 = RateFlowODE.from_eqx(rf.save_to_eqx(tmp_path, "rf"))
    assert np.array_equal(rf2.stoich_matrix, f32(rf.stoich_matrix)) and rf2.reaction_size == 3
    assert np.array_equal(rf2.mass_constraint, f32(rf.mass_constraint)) and abs(rf2.max_time - 7.5) < 1e-6
    assert all(np.array_equal(a, f32(b)) for a, b in zip(rf2.weights, rf.weights))
    assert rf2.biases[-1] is None and rf2.spec.final_activation == "relu"

    mech = ctx.Model.load(GOLD / "uode_model.json")
    for i, p in enumerate(mech.get_parameter_order()):
        mech.parameters[p].value = 0.05 + 0.01 * i
    uo = UniversalODE.from_model(mech, 8, 2, seed=4, max_time=3.0)
    uo2 = UniversalODE.from_eqx(uo.save_to_eqx(tmp_path, "uode.eqx"))
    assert np.array_equal(uo2.gate_matrix, f32(uo.gate_matrix))
    assert np.array_equal(uo2.alpha_residual, f32(uo.alpha_residual))
    assert np.array_equal(uo2.parameters, f32(uo.parameters))
    assert uo2.state_order == uo.state_order and uo2.model.get_parameter_order() == mech.get_parameter_order()
    assert [str(o.equation) for o in uo2.model.odes.values()] == [str(o.equation) for o in mech.odes.values()]

    rb = NeuralODE(2, 6, 2, ["x", "y"], activation=RBFLayer(0.4), key=3)
    rb2 = NeuralODE.from_eqx(rb.save_to_eqx(tmp_path, "rbf"))
    assert rb2.spec.rbf and abs(rb2.rbf_gamma - 0.4) < 1e-6
    assert all(np.array_equal(a, f32(b)) for a, b in zip(rb2.rbf_mu, rb.rbf_mu))
    # the reference's own checkpoint still loads
    ref = NeuralODE.from_eqx(GOLD / "menten_trained.eqx")
    assert ref.spec.activation == "celu" and [w.shape for w in ref.weights] == [(4, 2), (4, 4), (1, 4)]


def test_rbf_activation_in_the_kernel_pack_equals_the_pairwise_form():
    """The generated code evaluates exp(-(gW x^2 - 2g sum(mu) x + g sum(mu^2))); the reference sums
    (x - mu_j)^2 over the centres (rbf.py:32-39): identical up to rounding."""
    from catalax_b200.neural import NeuralODE, RBFLayer
    from catalax_b200.tools import codegen_mlp as cm
    from oracle import mlp_oracle as mo

    net = NeuralODE(3, 8, 2, ["a", "b", "c"], activation=RBFLayer(0.7), key=2)
    lay = cm.layout_of(net.spec)
    pack = cm.pack_parameters(net.spec, net.weights, net.biases, 1.0, np.float64, **net._pack_extra())
    x = np.random.default_rng(0).normal(size=(5, 8))
    for l in range(2):
        a, b, c = pack[lay.rbf_off[l]: lay.rbf_off[l] + 3]
        np.testing.assert_allclose(np.exp(-((a * x - b) * x + c)), mo.rbf_layer(x, net.rbf_mu[l], 0.7),
                                   rtol=1e-12)
    assert cm.tc_dims(net.spec) is None and cm.adjoint_dims(net.spec) is None
    assert "exp(-((th[" in cm.generate_mlp_field(net.spec).source


def test_model_json_round_trip(tmp_path):
    m = ctx.Model.load(GOLD / "uode_model.json")
    m.parameters[m.get_parameter_order()[0]].value = 0.25
    d = m.to_dict()
    assert set(d) >= {"name", "states", "odes", "parameters"}
    m2 = ctx.Model.from_dict(json.loads(json.dumps(d, default=str)))
    assert m2.get_state_order() == m.get_state_order()
    assert m2.get_parameter_order() == m.get_parameter_order()
    assert m2.parameters[m.get_parameter_order()[0]].value == 0.25
    r1 = m._replace_assignments(); r2 = m2._replace_assignments()
    assert [str(r1.odes[s].equation) for s in m.get_state_order()] == \
           [str(r2.odes[s].equation) for s in m.get_state_order()]
    assert ctx.Model.load(m.save(tmp_path, "roundtrip")).get_state_order() == m.get_state_order()


def test_sampler_transforms_and_entry_points():
    import torch

    from catalax_b200.mcmc import HMC, MCMCConfig, priors
    from catalax_b200.mcmc.sampler import _Transform, _transform_of

    u = torch.linspace(-3, 3, 13, dtype=torch.float64)
    for tf in (_transform_of(priors.Uniform(low=1e-6, high=200.0)), _Transform(0.0, None), _Transform()):
        x, lj, dx, dlj = tf.forward(u)
        np.testing.assert_allclose(tf.inverse(x).numpy(), u.numpy(), atol=1e-9)
        h = 1e-6
        xp, ljp, _, _ = tf.forward(u + h); xm, ljm, _, _ = tf.forward(u - h)
        np.testing.assert_allclose(((xp - xm) / (2 * h)).numpy(), dx.numpy(), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(((ljp - ljm) / (2 * h)).numpy(), dlj.numpy(), rtol=1e-6, atol=1e-8)
        # log|J| is the log of dx/du
        np.testing.assert_allclose(lj.numpy(), np.log(dx.numpy()), atol=1e-12)
    hmc = HMC(num_warmup=10, num_samples=5, num_chains=3, dt0=0.5)
    assert hmc.config.dt0 == 0.5 and hmc.config.max_steps == 64 ** 4 and hmc.config.target_accept_prob == 0.8
    assert HMC.from_config(MCMCConfig(1, 2)).config.num_samples == 2
    # mcmc.py:73-79: max_steps travels as t1; the solves keep SimulationConfig's defaults
    sc = hmc.config.to_simulation_config()
    assert sc.t1 == 64 ** 4 and sc.max_steps == 4096 and sc.rtol == 1e-5


def test_posterior_helper_functions():
    from catalax_b200.mcmc import loo, predictive

    rng = np.random.default_rng(0)
    post = {"K_m": rng.normal(size=(4, 50)), "v_max": rng.normal(size=(4, 50)),
            "sigma_y": np.abs(rng.normal(size=(4, 50, 1)))}
    th, (nc, nd) = predictive.stack_thetas(post, ["K_m", "v_max"])
    assert th.shape == (200, 2) and (nc, nd) == (4, 50) and th[51, 1] == post["v_max"][1, 1]
    assert predictive.sigma_y_draws(post, 4, 50, 3).shape == (200, 3)
    assert (predictive.sigma_y_draws({}, 4, 50, 2) == 1).all()
    sub = predictive.subsample(post, 7)
    assert sub["K_m"].shape == (4, 7) and sub["K_m"][2, -1] == post["K_m"][2, -1]
    assert predictive.subsample(post, None) is post
    t = predictive.ensemble_times(np.array([[0.0, 5.0, 10.0], [2.0, 3.0, 4.0]]), 5)
    np.testing.assert_allclose(t, [[0, 2.5, 5, 7.5, 10], [2, 2.5, 3, 3.5, 4]])
    var = predictive.aleatoric_variance(post, (200, 2, 3, 1))
    assert var.shape == (200, 2, 3, 1) and var[3, 1, 2, 0] == post["sigma_y"].reshape(200, 1)[3, 0] ** 2
    ll = rng.normal(size=(4, 50, 2, 3, 1))
    mask = np.ones((2, 3, 1), bool); mask[0, 0, 0] = False
    e = loo.pointwise_elpd_is(ll, mask)
    assert np.isnan(e[0, 0, 0]) and np.isfinite(e[1]).all()
    want = -np.log(np.mean(np.exp(-ll.reshape(200, 2, 3, 1)), axis=0))
    np.testing.assert_allclose(e[1], want[1], rtol=1e-12)


def test_workloads_are_shared_with_the_tests():
    from catalax_b200 import workloads as wl
    from tests import cases

    assert cases.michaelis_menten is wl.michaelis_menten and cases.mm_inputs is wl.mm_inputs
    y0, th, c = wl.mm_inputs(16, seed=3)
    assert (th[:, 0] >= 0.5).all() and (th[:, 0] <= 50).all()
    _, th2, _ = wl.mm_inputs(4096, seed=3, km_range=wl.KM_RANGE_SURVEY)
    assert (th2[:, 0] <= 1.0).all() and th2[:, 0].min() < 0.02
    import ast
    src = open(__import__("pathlib").Path(__file__).parents[1] / "bench.py").read()
    imports = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.ImportFrom)]
    assert not any((n.module or "").startswith("tests") for n in imports), "bench.py must not import tests/"


def test_batched_hmc_adapts_step_size_and_mass_on_a_badly_scaled_gaussian():
    """The sampler behind ``HMC.run`` on a density with a known answer (CPU, no engine): scales
    spanning 3000x must be recovered, with the acceptance rate near its target."""
    import torch

    from catalax_b200.mcmc.sampler import hmc_sample

    scales = torch.tensor([1.0, 0.01, 30.0], dtype=torch.float64)
    mean = torch.tensor([2.0, -1.0, 50.0], dtype=torch.float64)

    def lpg(u):
        z = (u - mean) / scales
        return -0.5 * (z * z).sum(1), -z / scales

    gen = torch.Generator().manual_seed(1)
    u = torch.rand((32, 3), generator=gen, dtype=torch.float64) * 4 - 2
    lp, g = lpg(u)
    keep, st = hmc_sample(lpg, u, lp, g, num_warmup=300, num_samples=300, max_steps=15, generator=gen)
    x = torch.stack(keep, 1).reshape(-1, 3)
    assert 0.6 < st["mean_accept_prob"] <= 1.0 and st["num_divergences"] == 0
    assert torch.allclose(x.mean(0), mean, atol=0.0, rtol=0.0) or \
        (((x.mean(0) - mean) / scales).abs() < 0.1).all()
    assert ((x.std(0) / scales - 1).abs() < 0.1).all()
