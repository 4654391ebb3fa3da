"""marl_b200 -- B200-native (sm_100a) hot path of Skylarking/MARL behind the reference's Python surface.

    from marl_b200.controller.share_params import SharedMAC
    from marl_b200.algorithm.q_learner import QLearner
    from marl_b200.network.mixer import VDNMixer, QMixMixer
    from marl_b200.env.single_state_matrix_game import TwoAgentsMatrixGame, BatchedMatrixGame

The compute lives in ``lib/libmarl_b200.so`` (hand-written CUDA behind the C ABI of
``include/marl_b200.h``), built in-tree by ``python -m marl_b200.build``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
