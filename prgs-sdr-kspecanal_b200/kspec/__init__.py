"""kspec — B200-native spectrum hot path for kSpecAnal (hanishkvc/prgs-sdr-kspecanal).

``kspec.engine.Plan``   object layer over the C ABI of libkspec.so (include/kspec.h)
``kspec.hotpath``       host-side mirror of the reference's functions (sdr_curscan, zero_span, _scan_range, ...)
``kspec.synth``         synthetic IQ + file/array backed RtlSdr stand-in (headless runs)

Importing the package does not load the CUDA library; the first Plan does, and fails loudly without it.
"""
__version__ = "0.1.0"
