"""GPU parity of the dense-layer primitives (nn.Linear and its autograd duals, network/q_network.py:17,20) against a
float64 torch reference, through both GEMM back ends: the register-staged tcgen05 kernels (csrc/linear.cu) and the
TMA-fed persistent tcgen05 kernel (csrc/tgemm.cu), plus a learner step on the TMA path against the oracle."""
import numpy as np
import pytest
import torch

from marl_b200 import _lib as L
from marl_b200.synthetic import synthetic_batch
from oracle import marl_oracle as MO
from tests import parity_util as PU

pytestmark = pytest.mark.gpu

TOL_GEMM = 3e-6        # 3xTF32: measured <= 1.8e-6 (max-norm relative) on every shape below; fp32 cuBLAS sits at ~1e-6


@pytest.fixture(params=[(0, 0), (0, 1), (1, 0)], ids=["linear.cu", "linear.cu-deterministic", "tgemm.cu"])
def backend(request):
    tg, det = request.param
    prev = L.load().marl_tgemm_enable(tg)
    prev_det = L.load().marl_set_deterministic(det)
    L.ensure_scratch()
    yield request.param
    L.load().marl_tgemm_enable(prev)
    L.load().marl_set_deterministic(prev_det)


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


SHAPES = [(19200, 64, 96, 96), (19200, 192, 64, 64), (3840, 256, 120, 120), (18, 64, 8, 8), (1000, 40, 36, 36), (300, 130, 100, 104),
          (5000, 11, 64, 64), (257, 64, 328, 328), (129, 33, 7, 12)]


@pytest.mark.parametrize("shape", SHAPES)
def test_linear_primitives_match_float64(backend, shape):
    M, N, K, ldx = shape
    torch.manual_seed(M + N + K)
    dev = "cuda"
    xs = torch.randn(M, ldx, device=dev)
    x = xs[:, :K]
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    L.call("marl_linear_fwd", xs.data_ptr(), ldx, w.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, N, K, 1, L.stream_ptr())
    ref = (x.double() @ w.double().t() + b.double()).clamp_min(0)
    assert _rel(y, ref) < TOL_GEMM
    dy = torch.randn(M, N, device=dev)
    dx = torch.empty(M, K, device=dev)
    L.call("marl_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, xs.data_ptr(), ldx, dx.data_ptr(), K, M, N, K, L.stream_ptr())
    assert _rel(dx, (dy.double() @ w.double()) * (x > 0)) < TOL_GEMM
    dw = torch.full((N, K), 0.5, device=dev)          # the weight gradient ACCUMULATES into dw / db
    db = torch.full((N,), -0.25, device=dev)
    L.call("marl_linear_wgrad", dy.data_ptr(), N, xs.data_ptr(), ldx, dw.data_ptr(), K, db.data_ptr(), M, N, K, L.stream_ptr())
    assert _rel(dw, dy.double().t() @ x.double() + 0.5) < TOL_GEMM
    assert _rel(db, dy.double().sum(0) - 0.25) < TOL_GEMM


def test_weight_gradient_is_bitwise_reproducible(backend):
    """With marl_set_deterministic(1) both back ends reduce the row splits in a fixed order (no atomics): repeated calls give
    identical bits (csrc/tgemm.cu always does)."""
    prev = L.load().marl_set_deterministic(1)
    try:
        _check_reproducible()
    finally:
        L.load().marl_set_deterministic(prev)


def _check_reproducible():
    M, N, K = 19200, 192, 64
    torch.manual_seed(0)
    dy, x = torch.randn(M, N, device="cuda"), torch.randn(M, K, device="cuda")
    outs = []
    for _ in range(3):
        dw, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
        L.call("marl_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, db.data_ptr(), M, N, K, L.stream_ptr())
        outs.append((dw.clone(), db.clone()))
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])


@pytest.mark.parametrize("alg", ["qmix", "vdn"])
def test_learner_step_on_tma_gemm_path_matches_oracle(alg):
    """The grouped TMA launches of the agent (composite input fill, shifted hidden operand, gathered dq operand,
    deterministic split reduce) behind QLearner.train: loss and clipped gradients vs the oracle at 1e-5."""
    prev = L.load().marl_tgemm_enable(1)
    try:
        args = PU.make_args(alg, 5, 11, 80, 120, 24)
        learner, st = PU.build_pair(args)
        batch = synthetic_batch(3, 16, 24, 5, 11, 80, 120)
        for step in range(2):
            loss = learner.train({k: v.copy() for k, v in batch.items()}, step)
            oloss, info = MO.train_step(st, batch, step)
            assert abs(loss - oloss) <= 1e-5 * abs(oloss), (step, loss, oloss)
            if step == 0:
                for g, k, p in st.flat_params():
                    mine = dict(PU.module_groups(learner)[g].named_parameters())[k]
                    assert PU.rel_err(mine.grad, info["clipped_grads"][f"{g}.{k}"]) < 1e-5, (g, k)
    finally:
        L.load().marl_tgemm_enable(prev)


# ------------------------------------------------------------------------------------------ fused input layers (csrc/front.cu)
@pytest.mark.parametrize("shape", [
    (32, 120, 5, 11, 80),      # BASELINE config 2: 19 200 rows, I = 96
    (3, 7, 2, 3, 4),           # tiny: one ragged tile, I = 9 (one k-tile, all of it composed or clipped)
    (5, 33, 3, 5, 12),         # rows not a multiple of 128, I = 20
    (2, 9, 8, 14, 128),        # 3s5z width: I = 150 (ten k-tiles, the last one partial)
    (4, 6, 2, 6, 248),         # I = 256: the widest the fused kernel takes
])
def test_fused_input_layers_match_the_separate_launches_and_float64(shape):
    """marl_agent_unroll_fwd with the fused front kernel (TMA in, A operands in tensor memory, TMA out) against the same entry
    point with the layers launched one by one (marl_front_enable(0)) and against float64: x = relu(fc1([obs | last action | id])),
    gi = W_ih x + b_ih (network/q_network.py:17-19), then the hidden states and q both paths feed."""
    import ctypes as C
    from marl_b200 import _lib as L
    from marl_b200.network.q_network import AGENT_FLAT_ORDER, agent_param_struct
    B, T, N, A, O = shape
    I, H, rows = O + A + N, 64, B * T * N
    torch.manual_seed(B * 1000 + T)
    dev = "cuda"
    obs = torch.randn(B, T, N, O, device=dev)
    onehot = torch.zeros(B, T, N, A, device=dev)
    onehot.scatter_(3, torch.randint(0, A, (B, T, N, 1), device=dev), 1.0)
    sizes = {"fc1_w": (H, I), "fc1_b": (H,), "w_ih": (3 * H, H), "w_hh": (3 * H, H), "b_ih": (3 * H,), "b_hh": (3 * H,),
             "fc2_w": (A, H), "fc2_b": (A,)}
    flat = torch.empty(sum(int(np.prod(v)) + 3 & ~3 for v in sizes.values()), device=dev)     # 16-byte aligned parameters
    params, off = {}, 0
    for k in L.AGENT_KEYS:
        n = int(np.prod(sizes[k]))
        params[k] = flat[off:off + n].view(sizes[k]).normal_(0, 0.3)
        off += n + 3 & ~3
    named = {name: params[k].data_ptr() for k, name in zip(L.AGENT_KEYS, AGENT_FLAT_ORDER)}

    def run(fused):
        prev = L.load().marl_front_enable(1 if fused else 0)
        try:
            out = {k: torch.full(s, float("nan"), device=dev) for k, s in
                   (("q", (rows, A)), ("hidden", (rows, H)), ("x", (rows, H)), ("gi", (rows, 3 * H)), ("gates", (rows, 4 * H)))}
            st = L.UnrollStream()
            st.obs, st.onehot, st.shift_onehot, st.full_input, st.h0_from, st.h0 = obs.data_ptr(), onehot.data_ptr(), 1, 0, -1, None
            st.params = agent_param_struct(named)
            st.q, st.hidden, st.h_last = out["q"].data_ptr(), out["hidden"].data_ptr(), None
            st.x, st.gi, st.gates = out["x"].data_ptr(), out["gi"].data_ptr(), out["gates"].data_ptr()
            d = L.Dims(B, T, N, A, O, 0)
            L.profile(True)
            L.call("marl_agent_unroll_fwd", C.byref(d), C.byref(st), 1, L.stream_ptr())
            launched = L.profile_collect()
            L.profile(False)
            assert ("agent_front_kernel" in launched) == fused, launched       # the path under test really ran
            return out
        finally:
            L.load().marl_front_enable(prev)

    fused, plain = run(True), run(False)
    last = torch.cat([torch.zeros(B, 1, N, A, device=dev), onehot[:, :-1]], 1)                  # share_params.py:92-101
    ident = torch.eye(N, device=dev).expand(B, T, N, N)
    inp = torch.cat([obs, last, ident], 3).reshape(rows, I).double()
    x64 = torch.relu(inp @ params["fc1_w"].double().t() + params["fc1_b"].double())
    gi64 = x64 @ params["w_ih"].double().t() + params["b_ih"].double()
    for name, ref in (("x", x64), ("gi", gi64)):
        for tag, got in (("fused", fused[name]), ("separate", plain[name])):
            err = float((got.double() - ref).abs().max() / ref.abs().max())
            assert err < 3e-6, (name, tag, err)
    for name in ("hidden", "q"):
        err = float((fused[name] - plain[name]).abs().max() / plain[name].abs().max())
        assert err < 3e-6, (name, err)
