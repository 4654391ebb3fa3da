"""The CPU oracle against the goldens produced by the unmodified reference (CPU, no GPU)."""
import numpy as np
import pytest
import torch

from oracle import marl_oracle as MO
from tests import golden_util as GU

TOL = 2e-5   # fp32, same torch ops in a different composition


@pytest.mark.parametrize("name", GU.learner_cases())
def test_oracle_reproduces_reference_training(name):
    z = GU.load(name)
    cfg = GU.cfg_from(z)
    st = MO.LearnerState(cfg, GU.init_params(z))
    batch = GU.batch_of(z)
    losses = []
    for step in range(int(z["meta/n_steps"])):
        loss, info = MO.train_step(st, batch, step)
        losses.append(loss)
        if step == 0:
            assert info["L"] == int(z["step0/L"])
            assert GU.rel_err(info["q_evals"], z["step0/q_evals"]) < TOL
            assert GU.rel_err(info["hidden_evals"], z["step0/hidden_evals"]) < TOL
            assert GU.rel_err(info["q_targets"], z["step0/q_targets"]) < TOL
            if "step0/cur_max_actions" in z:
                assert np.array_equal(info["a_star"].squeeze(3).numpy(), z["step0/cur_max_actions"])
            if "step0/q_tot" in z:
                assert GU.rel_err(info["q_tot"], z["step0/q_tot"]) < TOL
            if "step0/hidden_targets" in z:
                assert GU.rel_err(info["hidden_targets"], z["step0/hidden_targets"]) < TOL
            for k, g in info["clipped_grads"].items():
                key = "clipped_grad/" + k.replace(".", "/", 1)
                if g is None:
                    assert key not in z
                else:
                    assert GU.rel_err(g, z[key]) < 5e-5, k
    assert np.allclose(losses, z["loss"], rtol=TOL, atol=0), (losses, z["loss"])
    for g in ("agent", "mixer") + (("v",) if cfg.alg == "qtran_base" else ()):
        for k, v in GU.group(z, "final/" + g).items():
            assert GU.rel_err(st.params[g][k].detach(), v) < 5e-5, (g, k)
    for g in ("agent", "mixer"):
        for k, v in GU.group(z, "final_target/" + g).items():
            assert GU.rel_err(st.target[g][k], v) < 5e-5, (g, k)


def test_oracle_fp64_close_to_fp32():
    z = GU.load("tiny_qmix_rms")
    cfg = GU.cfg_from(z)
    a = MO.LearnerState(cfg, GU.init_params(z))
    b = MO.LearnerState(cfg, GU.init_params(z), dtype=torch.float64)
    la, ia = MO.train_step(a, GU.batch_of(z), 0)
    lb, ib = MO.train_step(b, GU.batch_of(z), 0)
    assert abs(la - lb) / abs(lb) < 1e-5
    assert GU.rel_err(ia["q_tot"], ib["q_tot"]) < 1e-5


def test_gru_explicit_matches_aten():
    torch.manual_seed(1)
    x, h = torch.randn(7, 64), torch.randn(7, 64)
    g = torch.nn.GRUCell(64, 64)
    ref = g(x, h)
    mine = MO.gru_cell_explicit(x, h, g.weight_ih, g.weight_hh, g.bias_ih, g.bias_hh)
    assert torch.allclose(ref, mine, atol=1e-6)


def test_max_episode_len_semantics():
    term = np.zeros((3, 5, 1))
    assert MO.max_episode_len(term, 5) == 5           # nobody terminated -> limit (q_learner.py:59-60)
    term[0, 2:, 0] = 1
    term[1, 1:, 0] = 1
    assert MO.max_episode_len(term, 5) == 3
