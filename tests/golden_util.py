"""Helpers to load tests/golden/*.npz fixtures (made by oracle/make_golden.py)."""
import glob
import os

import numpy as np
import torch

from oracle import marl_oracle as MO

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


# fixtures with their own layout (tests of their own): not full learner cases
OTHER_FIXTURES = ("matrix_game_env", "choose_action_3s5z", "checkpoint_losses", "qmix_2s3z_seeds", "rollout_multistep", "separated_")


def learner_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(OTHER_FIXTURES))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def group(z, prefix):
    """{'init/agent/fc1.weight': arr} -> {'fc1.weight': tensor} for prefix 'init/agent'."""
    pl = prefix + "/"
    return {k[len(pl):]: torch.from_numpy(np.array(v)) for k, v in z.items() if k.startswith(pl)}


def cfg_from(z):
    N, A, O, S, T = (int(x) for x in z["meta/dims"])
    return MO.make_cfg(alg=str(z["meta/alg"]), optimizer=str(z["meta/optimizer"]), n_agents=N, n_actions=A,
                       obs_shape=O, state_shape=S, episode_limit=T, double_q=bool(int(z["meta/double_q"])),
                       lr=float(z["meta/lr"]), target_update_cycle=int(z["meta/target_update_cycle"]),
                       num_kernel=int(z["meta/num_kernel"]), adv_hypernet_embed=int(z["meta/adv_hypernet_embed"]),
                       hypernet_embed=int(z["meta/hypernet_embed"]), qtran_hidden_dim=int(z["meta/qtran_hidden_dim"]),
                       two_hyper_layers=bool(int(z["meta/two_hyper_layers"])) if "meta/two_hyper_layers" in z else False,
                       hyper_hidden_dim=int(z["meta/hyper_hidden_dim"]) if "meta/hyper_hidden_dim" in z else 64,
                       adv_hypernet_layers=int(z["meta/adv_hypernet_layers"]) if "meta/adv_hypernet_layers" in z else 3)


def init_params(z):
    groups = ["agent", "mixer"] + (["v", "q_sum_mixer"] if str(z["meta/alg"]) == "qtran_base" else [])
    return {g: group(z, "init/" + g) for g in groups}


def batch_of(z):
    return {k[len("batch/"):]: np.array(v) for k, v in z.items() if k.startswith("batch/")}


def rel_err(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(x - y)) / max(float(np.max(np.abs(y))), 1e-30)) if x.size else 0.0
