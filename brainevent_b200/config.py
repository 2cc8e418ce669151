"""Backend selection flags (ref: brainevent/config.py:220 set_backend / get_backend).

Only one backend exists here: 'b200' (hand-written sm_100a CUDA behind the C ABI) on platform
'gpu'.  The functions are kept so code written against the reference keeps working; asking for
any other backend raises instead of silently falling back.
"""
_DEFAULT = {"gpu": "b200"}
BACKENDS = ("b200",)


def set_backend(platform: str, backend):
    if platform != "gpu":
        raise ValueError(f"brainevent_b200 only runs on platform 'gpu', got {platform!r}")
    if backend is not None and backend not in BACKENDS:
        raise ValueError(f"backend {backend!r} is not available; brainevent_b200 provides {BACKENDS}")
    _DEFAULT[platform] = backend or "b200"


def get_backend(platform: str = "gpu"):
    return _DEFAULT.get(platform)
