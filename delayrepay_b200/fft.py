"""delayrepay.fft entry point  (reference fft.py:9-12).

The reference's ``fft(self, *args)`` carries a stray ``self`` parameter and is broken on its
CPU backend; the signature here is NumPy's.  A cuFFT-backed
implementation behind the C ABI is the next row; until then the
entry point exists and raises, rather than silently computing on the host.
"""


def fft(a, n=None, axis=-1, norm=None):
    raise NotImplementedError(
        "delayrepay_b200.fft.fft: cuFFT binding not built yet (SURVEY.md section 8f rank 3); "
        "there is no CPU fallback")
