"""ppasr_b200: the PPASR inference hot path (fbank -> encoder -> CTC projection -> greedy / beam search) native on NVIDIA B200.

Same package-level names as ppasr/__init__.py. Nothing is imported eagerly: the CUDA library is loaded on first use."""
__version__ = "0.1.0"
# models the engine implements (ppasr/__init__.py:3)
SUPPORT_MODEL = ['squeezeformer', 'efficient_conformer', 'conformer', 'deepspeech2']
