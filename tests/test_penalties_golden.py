"""Training-loss regularisers against values produced by the reference's own penalties package
(tests/golden/make_penalties_golden.py executes catalax/neural/penalties/*.py unmodified over a
numpy stand-in for jax.numpy): every penalty function, the three presets, ``update_alpha``; the
host autograd gradients against finite differences of the reference functions."""
import types
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
G = np.load(Path(__file__).parent / "golden" / "penalties_reference_lines.npz")


def _mlp():
    return dict(weights=[G["w0"].copy(), G["w2"].copy()], biases=[G["w1"].copy(), None], rbf_mu=None)


def _rateflow(mass=True):
    return types.SimpleNamespace(kind="rateflow", stoich_matrix=G["rf_stoich"].copy(),
                                 mass_constraint=G["rf_mass"].copy() if mass else None, **_mlp())


def _uode():
    return types.SimpleNamespace(kind="universalode", gate_matrix=G["uo_gate"].copy(),
                                 alpha_residual=G["uo_alpha"].copy(), parameters=G["uo_parameters"].copy(),
                                 **_mlp())


RF = {"density": "penalize_density", "bipolar": "penalize_non_bipolar",
      "conservation": "penalize_non_conservative", "duplicate_reactions": "penalize_duplicate_reactions",
      "integer": "penalize_non_integer", "sparsity": "l1_stoich_penalty", "null_space": "penalize_null_space",
      "l2": "l2_regularisation", "l1": "l1_regularisation"}
UO = {"l2_gate": "l2_reg_gate", "l1_gate": "l1_reg_gate", "l2_residual": "l2_reg_alpha",
      "l1_residual": "l1_reg_alpha", "l2_mlp": "l2_regularisation", "l1_mlp": "l1_regularisation"}


@pytest.mark.parametrize("name", list(RF))
def test_rateflow_penalty_functions(name):
    from catalax_b200.neural import penalties as pn

    pen = pn.Penalties().add_penalty(name=name, fun=getattr(pn, RF[name]), alpha=0.37)
    val, grads = pen.value_and_grads(_rateflow())
    np.testing.assert_allclose(val, G[f"rf/{name}/value"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(grads["stoich"][0], G[f"rf/{name}/grad_stoich"], rtol=2e-6, atol=2e-8)


@pytest.mark.parametrize("name", list(UO))
def test_universal_ode_penalty_functions(name):
    from catalax_b200.neural import penalties as pn

    pen = pn.Penalties().add_penalty(name=name, fun=getattr(pn, UO[name]), alpha=0.21)
    val, grads = pen.value_and_grads(_uode())
    np.testing.assert_allclose(val, G[f"uo/{name}/value"], rtol=1e-12)
    np.testing.assert_allclose(grads["gate"][0], G[f"uo/{name}/grad_gate"], rtol=2e-6, atol=2e-8)
    np.testing.assert_allclose(grads["alpha"][0], G[f"uo/{name}/grad_alpha"], rtol=2e-6, atol=2e-8)


def test_presets_hold_the_reference_penalties_in_the_reference_order():
    from catalax_b200.neural.penalties import Penalties

    all_kw = dict(alpha=0.2, density_alpha=0.3, bipolar_alpha=0.05, conservation_alpha=0.4,
                  duplicate_reactions_alpha=1.5, integer_alpha=0.15, sparsity_alpha=0.01, l2_alpha=1e-3,
                  null_space_alpha=0.6)
    for name, kw in {"default": {}, "all": all_kw}.items():
        pen = Penalties.for_rateflow(**kw)
        assert [p.name for p in pen.penalties] == list(G[f"rf/preset_{name}/names"])
        val, grads = pen.value_and_grads(_rateflow())
        np.testing.assert_allclose(val, G[f"rf/preset_{name}/value"], rtol=1e-12)
        np.testing.assert_allclose(grads["stoich"][0], G[f"rf/preset_{name}/grad_stoich"], rtol=2e-6, atol=2e-8)
        assert pen(_rateflow()) == pytest.approx(val, rel=1e-15)
    upd = Penalties.for_rateflow().update_alpha(0.05, l2=0.5)
    np.testing.assert_allclose(upd(_rateflow()), G["rf/update_alpha/value"], rtol=1e-12)
    for name, kw in {"default": {}, "l1": dict(l1_gate_alpha=0.02, l1_residual_alpha=0.03, l1_mlp_alpha=0.004,
                                               l2_gate_alpha=None)}.items():
        pen = Penalties.for_universal_ode(**kw)
        assert [p.name for p in pen.penalties] == list(G[f"uo/preset_{name}/names"])
        np.testing.assert_allclose(pen(_uode()), G[f"uo/preset_{name}/value"], rtol=1e-12)
    node = types.SimpleNamespace(kind="neuralode", **_mlp())
    for name, kw in {"default": {}, "both": dict(l2_alpha=2e-3, l1_alpha=5e-4)}.items():
        pen = Penalties.for_neural_ode(**kw)
        assert [p.name for p in pen.penalties] == list(G[f"no/preset_{name}/names"])
        val, grads = pen.value_and_grads(node)
        np.testing.assert_allclose(val, G[f"no/preset_{name}/value"], rtol=1e-12)
        np.testing.assert_allclose(grads["weights"][0], G[f"no/preset_{name}/grad_w0"], rtol=2e-6, atol=2e-8)
        assert [g.shape for g in grads["weights"]] == [G["w0"].shape, G["w1"].shape, G["w2"].shape]


def test_null_space_without_a_mass_constraint_is_zero_and_wrong_models_are_rejected():
    from catalax_b200.neural import penalties as pn

    pen = pn.Penalties().add_penalty("null_space", pn.penalize_null_space, 0.37)
    assert pen(_rateflow(mass=False)) == float(G["rf/null_space_without_constraint"]) == 0.0
    with pytest.raises(AssertionError, match="RateFlowODE"):
        pn.Penalties().add_penalty("d", pn.penalize_density, 0.1)(_uode())
    with pytest.raises(AssertionError, match="UniversalODE"):
        pn.Penalties().add_penalty("g", pn.l2_reg_gate, 0.1)(_rateflow())
    with pytest.raises(ValueError, match="must accept 'model' and 'alpha'"):
        pn.Penalties().add_penalty("bad", lambda x: x, 0.1)


def test_custom_penalty_and_all_array_leaves():
    """A user function gets the torch view; l2 runs over EVERY array leaf (neuralbase.py:576-584),
    max_time included once training has turned it into an array (trainer.py:85-89)."""
    from catalax_b200.neural import NeuralODE, penalties as pn

    net = NeuralODE(3, 8, 2, ["a", "b", "c"], key=0)
    base = pn.Penalties.for_neural_ode(l2_alpha=1.0)(net)
    assert base == pytest.approx(sum(float((a ** 2).sum()) for a in net.get_weights_and_biases()))
    net.max_time, net._max_time_is_leaf = 7.0, True
    assert pn.Penalties.for_neural_ode(l2_alpha=1.0)(net) == pytest.approx(base + 49.0)
    first = pn.Penalties().add_penalty("first", lambda model, alpha: alpha * model.get_weights_and_biases()[0].sum(), 2.0)
    val, grads = first.value_and_grads(net)
    assert val == pytest.approx(2.0 * net.weights[0].sum())
    assert np.allclose(grads["weights"][0], 2.0) and all(np.all(g == 0) for g in grads["weights"][1:])


@pytest.mark.parametrize("loss", ["l2", "l1"])
def test_masked_training_loss_matches_the_reference_grad_loss(loss):
    """trainer.py:270-311 (`grad_loss`, executed unmodified by the golden maker): per-point loss,
    temporal-dropout mask, normalisation by the kept points, plus the penalties."""
    from catalax_b200.neural.neuralbase import masked_loss
    from catalax_b200.neural.penalties import Penalties

    pred, target, mask = (torch.as_tensor(G[f"loss/{loss}/{k}"]) for k in ("pred", "target", "mask"))
    data_term = float(masked_loss(pred, target, loss, mask))
    node = types.SimpleNamespace(kind="neuralode", **_mlp())
    penalty = Penalties.for_neural_ode(l2_alpha=1e-3)(node)
    np.testing.assert_allclose(penalty, G[f"loss/{loss}/penalty"], rtol=1e-12)
    np.testing.assert_allclose(data_term + penalty, G[f"loss/{loss}/total"], rtol=1e-12)
    # a callable loss and no mask
    custom = float(masked_loss(pred, target, lambda p, y: (p - y) ** 2))
    np.testing.assert_allclose(custom, float(((pred - target) ** 2).mean()), rtol=1e-12)
