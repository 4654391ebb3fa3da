"""Hyper-parameter namespace with the attribute names the reference's classes read.

The reference injects these into an argparse namespace (``common/arguments.py:9-45`` for the
command line, ``:86-147`` ``get_mixer_args`` for VDN/QMIX/QPLEX/QTRAN); the values below are its
defaults.  ``get_mixer_args`` accepts any namespace and fills the same attributes, so existing
entry scripts keep working; ``default_args`` builds one without argparse.
"""
from types import SimpleNamespace

COMMON_DEFAULTS = dict(          # common/arguments.py:12-44
    RTW=False, env='smac', map='2s3z', seed=123, alg='qmix', n_steps=800000, n_episodes=1,
    last_action=True, reuse_network=True, gamma=0.99, optimizer="RMS", evaluate_cycle=5000,
    evaluate_epoch=0, model_dir='./model', result_dir='./result', load_model=False, evaluate=False,
    cuda=True, replay_dir='',
)

MIXER_DEFAULTS = dict(           # common/arguments.py:86-147
    rnn_hidden_dim=64, qmix_hidden_dim=32, two_hyper_layers=False, hyper_hidden_dim=64, qtran_hidden_dim=64,
    lr=5e-4, epsilon=1, min_epsilon=0.05, epsilon_anneal_scale='step', train_steps=1, batch_size=32,
    buffer_size=int(5e3), save_cycle=5000, target_update_cycle=200, lambda_opt=1, lambda_nopt=1,
    grad_norm_clip=10, noise_dim=16, lambda_mi=0.001, lambda_ql=1, entropy_coefficient=0.001,
    adv_hypernet_embed=64, num_kernel=10, adv_hypernet_layers=3, weighted_head=True, hypernet_embed=64,
    is_minus_one=True, mixing_embed_dim=32, double_q=True,
)


def get_mixer_args(args):
    for k, v in MIXER_DEFAULTS.items():
        setattr(args, k, v)
    args.anneal_epsilon = (args.epsilon - args.min_epsilon) / 50000
    return args


def default_args(**overrides):
    """Namespace with the reference's defaults; env-derived fields (n_agents, n_actions, obs_shape,
    state_shape, episode_limit) come from ``env.get_env_info()`` as in main.py:24-29."""
    args = SimpleNamespace(**COMMON_DEFAULTS)
    get_mixer_args(args)
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def apply_env_info(args, env_info):
    args.n_actions = env_info["n_actions"]
    args.n_agents = env_info["n_agents"]
    args.state_shape = env_info["state_shape"]
    args.obs_shape = env_info["obs_shape"]
    args.episode_limit = env_info["episode_limit"]
    return args
