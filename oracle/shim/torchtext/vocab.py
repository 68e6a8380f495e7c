"""TEST INFRASTRUCTURE ONLY - stub; GloVe text loading is outside the hot path."""


def __getattr__(name):
    raise AttributeError(f"torchtext.vocab.{name} is not available in the oracle shim")
