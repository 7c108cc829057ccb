"""B200-native FCAM hot path of Text-Guided Face Recognition (TGFR).

`models/` mirrors the reference's `models/{attention,losses,metrics,magface}.py` symbols
(same names, signatures and error behaviour); `ops` holds the autograd bindings of the
C-ABI CUDA library `libtgfr_b200.so` (include/tgfr_b200.h); `distributed` shards the path
over one NVSwitch box.  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"
