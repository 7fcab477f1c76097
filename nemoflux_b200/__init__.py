"""nemoflux_b200 -- B200-native transect-flux hot path of pletzer/nemoflux.

``nemoflux_gpu``  mint-style Grid / PolylineIntegral + the batched flux series (CUDA via ctypes)
``field``         Field mirror of nemoflux/field.py backed by the GPU path
``datagen`` ``fluxviz`` ``fluxplot`` ``fluxexact``  drop-in command line tools (same flags)
"""
__version__ = '0.1.0'
