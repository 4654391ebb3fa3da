"""GPU parity of the acting surface (SURVEY 8(a) a19 and the north star's mac.forward alias): SharedMAC.choose_action /
SharedMAC.forward against the reference-pinned goldens, marl_epsgreedy_select with availability masks."""
import numpy as np
import pytest
import torch

from marl_b200 import _lib as L
from oracle import marl_oracle as MO
from tests import golden_util as GU
from tests import parity_util as PU

pytestmark = pytest.mark.gpu


def test_choose_action_matches_reference_golden():
    """share_params.py:37-72 driven like rollout.py:60-76 on the reference's literal 3s5z observations and availability
    masks (test_file/choose_action_test.py:7-208): greedy AND exploring steps under the golden's numpy seed.  Actions
    bit-exact (the hard requirement of a19), carried hidden state 1e-5."""
    from marl_b200.controller.share_params import SharedMAC
    z = GU.load("choose_action_3s5z")
    N, A, O = (int(x) for x in z["meta/dims"])
    args = PU.make_args("qmix", N, A, O, 216, 150)
    torch.manual_seed(123)                       # different init on purpose: the weights come from the golden
    mac = SharedMAC(args)
    mac.agent.load_state_dict({k: v for k, v in GU.group(z, "init/agent").items()})
    np.random.seed(int(z["meta/seed"]))
    mac.init_hidden(1)
    last = np.zeros((N, A))
    T = z["obs"].shape[0]
    for t in range(T):
        for a in range(N):
            act = int(mac.choose_action(z["obs"][t, a], last[a], a, z["avail"][t, a], float(z["eps"][t])))
            assert act == int(z["actions"][t, a]), (t, a)
            assert z["avail"][t, a, act] == 1
            last[a] = np.eye(A)[act]
        assert PU.rel_err(mac.hidden_states[0], z["hidden"][t]) < 1e-5, t


@pytest.mark.parametrize("name", ["tiny_qmix_rms", "ragged_qmix_rms"])
def test_mac_forward_alias_steps_through_the_episode(name):
    """mac.forward(ep_batch, t) for t = 0..L-1 (the north star's pymarl spelling) reproduces the reference's
    get_current_q_values (share_params.py:125-146) step by step, including the carried hidden state."""
    from marl_b200.controller.share_params import SharedMAC
    z = GU.load(name)
    N, A, O, S, T = (int(x) for x in z["meta/dims"])
    args = PU.make_args(str(z["meta/alg"]), N, A, O, S, T)
    mac = SharedMAC(args)
    mac.agent.load_state_dict(GU.group(z, "init/agent"))
    batch = GU.batch_of(z)
    Lq = int(z["step0/L"])
    B = batch["o"].shape[0]
    mac.init_hidden(B)
    ep = {"o": torch.as_tensor(batch["o"], dtype=torch.float32, device="cuda"),
          "u_onehot": torch.as_tensor(batch["u_onehot"], dtype=torch.float32, device="cuda")}
    with torch.no_grad():
        for t in range(Lq):
            q = mac.forward(ep, t)
            assert tuple(q.shape) == (B, N, A)
            assert PU.rel_err(q, z["step0/q_evals"][:, t]) < 1e-5, t
    assert PU.rel_err(mac.hidden_states.reshape(B, N, -1), z["step0/hidden_evals"][:, Lq - 1]) < 1e-5


@pytest.mark.parametrize("rows,A", [(1, 3), (40, 14), (4097, 11), (1000, 36)])
def test_epsgreedy_select_with_availability_masks(rows, A):
    """marl_epsgreedy_select against share_params.py:62-72 written in numpy: -inf mask, FIRST maximum (ties included),
    the host-drawn random action wins when exploring; one-hot output."""
    rng = np.random.RandomState(rows + A)
    q = rng.randn(rows, A).astype(np.float32)
    q[rng.rand(rows, A) < 0.2] = 0.5                                   # plenty of exact ties
    avail = (rng.rand(rows, A) < 0.6).astype(np.float32)
    avail[np.arange(rows), rng.randint(0, A, rows)] = 1.0            # at least one available action per row
    explore = (rng.rand(rows) < 0.3).astype(np.uint8)
    rand_a = np.array([rng.choice(np.nonzero(avail[i])[0]) for i in range(rows)], dtype=np.int64)
    masked = np.where(avail == 0, -np.inf, q)
    want = np.where(explore == 1, rand_a, masked.argmax(axis=1))
    dev = "cuda"
    act = torch.empty(rows, dtype=torch.int64, device=dev)
    onehot = torch.empty(rows, A, dtype=torch.float32, device=dev)
    tq, ta = torch.as_tensor(q, device=dev), torch.as_tensor(avail, device=dev)
    te, tr = torch.as_tensor(explore, device=dev), torch.as_tensor(rand_a, device=dev)
    L.call("marl_epsgreedy_select", rows, A, tq.data_ptr(), ta.data_ptr(), te.data_ptr(), tr.data_ptr(), act.data_ptr(),
           onehot.data_ptr(), L.stream_ptr())
    assert np.array_equal(act.cpu().numpy(), want)
    assert np.array_equal(onehot.cpu().numpy(), np.eye(A, dtype=np.float32)[want])
    # greedy only, no explore arrays
    L.call("marl_epsgreedy_select", rows, A, tq.data_ptr(), ta.data_ptr(), None, None, act.data_ptr(), None, L.stream_ptr())
    assert np.array_equal(act.cpu().numpy(), masked.argmax(axis=1))
