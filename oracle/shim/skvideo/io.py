def vwrite(*a, **k):
    raise RuntimeError("skvideo stub (oracle shim): video export is outside the hot path")
