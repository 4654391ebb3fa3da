"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference for the reference arm.

The reference (Skylarking/MARL) is pure Python with no build system, so ``pip install --target`` does not
apply.  This recipe copies the few source files of the hot path, byte for byte, from the
read-only checkout into ``oracle/_ref/`` -- git-ignored (never part of the history) but NOT gpurun-ignored,
so it travels to the GPU box like a built ``.so``.  ``__graft_entry__.build()`` runs it whenever
``/root/reference`` is mounted; on the GPU box the staged copy is simply used.

    python oracle/fetch_ref.py            # stage;  prints the file list with sha256

``load()`` imports the staged reference with the two shims of SURVEY.md appendix B (gym stub, np.float /
np.long) and returns its public classes.  Nothing under ``marl_b200/`` imports this module.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MARL_REFERENCE", "/root/reference")

# the files SURVEY.md section 8(c) lists for the path (+ what they import at module level)
FILES = (
    "algorithm/q_learner.py", "algorithm/qtran_learner.py", "controller/share_params.py",
    "network/q_network.py", "network/mixer.py", "network/RTW.py", "network/world_model.py",
    "common/arguments.py", "common/replaybuffer.py", "env/single_state_matrix_game.py", "rollout.py",
)


def stage(verbose=False):
    """Copy FILES from the reference checkout to oracle/_ref/ (no-op when the checkout is absent)."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    listing = []
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        data = open(s, "rb").read()
        if not (os.path.exists(d) and open(d, "rb").read() == data):
            if os.path.exists(d):
                os.chmod(d, 0o644)
            shutil.copyfile(s, d)
        listing.append((rel, hashlib.sha256(data).hexdigest()[:16]))
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write(f"# unmodified copies from {SRC} (oracle/fetch_ref.py)\n")
        for rel, h in listing:
            f.write(f"{h}  {rel}\n")
    if verbose:
        for rel, h in listing:
            print(h, rel)
    return True


def available():
    return os.path.exists(os.path.join(DST, "algorithm", "q_learner.py"))


def load():
    """Import the staged reference; returns a namespace with its public classes and arg helper."""
    import numpy as np
    if not available():
        raise RuntimeError("oracle/_ref is not staged (run python oracle/fetch_ref.py where /root/reference exists)")
    if DST not in sys.path:
        sys.path.insert(0, DST)
    sys.modules.setdefault("gym", types.SimpleNamespace(Env=object))       # env/single_state_matrix_game.py:3,123
    if not hasattr(np, "float"):
        np.float = float                                                    # env/...:7,84 ; algorithm/q_learner.py:219
    if not hasattr(np, "long"):
        np.long = np.int64
    from common.arguments import get_mixer_args
    from controller.share_params import SharedMAC
    from algorithm.q_learner import QLearner
    from algorithm.qtran_learner import QTRANLearner
    from common.replaybuffer import ReplayBuffer
    from env.single_state_matrix_game import TwoAgentsMatrixGame
    return types.SimpleNamespace(get_mixer_args=get_mixer_args, SharedMAC=SharedMAC, QLearner=QLearner,
                                 QTRANLearner=QTRANLearner, ReplayBuffer=ReplayBuffer,
                                 TwoAgentsMatrixGame=TwoAgentsMatrixGame)


def make_args(ns, alg, N, A, O, S, T, cuda=False, **kw):
    """Hand-built args (get_common_args() parses sys.argv): SURVEY.md appendix B."""
    a = types.SimpleNamespace(RTW=False, alg=alg, map="synthetic", last_action=True, reuse_network=True, gamma=0.99,
                              optimizer="RMS", model_dir="/tmp/marl_ref_model", result_dir="/tmp", cuda=cuda,
                              load_model=False, evaluate=False, evaluate_epoch=0, replay_dir="", n_episodes=1)
    ns.get_mixer_args(a)
    a.n_agents, a.n_actions, a.obs_shape, a.state_shape, a.episode_limit = N, A, O, S, T
    for k, v in kw.items():
        setattr(a, k, v)
    return a


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("staged" if ok else "reference checkout not found", DST)
