"""multimodaltraj_2_b200 -- B200-native (sm_100a) implementation of the multimodaltraj forecasting
hot path behind the reference's model / cell / train / sample call signatures.

Layout: ``csrc/`` hand-written CUDA kernels + the C-ABI (``include/mmt.h``), ``_lib.py`` ctypes
binding, ``ops.py`` tensor-level wrappers, and the host-side mirror of the reference interface
(``models/``, ``relational_inf_models/``, ``helper.py``, ``sample.py``, ``load_traj.py``,
``networkx_graph.py``, ``train.py``, ``argParser.py``).
"""
__version__ = "0.1.0"
