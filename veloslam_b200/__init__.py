"""B200-native Velodyne ingest hot path (VeloSLAM HDLParser/TransformManager drop-in).

The product is the C-ABI shared library built from ``veloslam_b200/csrc`` (declared in
``include/veloslam_b200.h``); this Python package is the thin harness over it
(ctypes bindings, synthetic inputs, pcap framing) used by tests and bench.py.
"""
__version__ = "0.1.0"
