"""Video writer mirror (utils/vis.py:142-161): formats, file names, joblib
compatibility with scripts/eval_owl_vit.py:52 (`joblib.load`), async writer ordering and error propagation."""
import os

import numpy as np
import pytest


def _frames(seed=0, f=6, h=16, w=24):
    rng = np.random.default_rng(seed)
    return rng.random((f, h, w, 3), dtype=np.float32)


def test_save_frames_formats_roundtrip(tmp_path):
    import joblib
    from PIL import Image
    from lvd_b200.utils.vis import save_frames
    fr = _frames()
    p = str(tmp_path / "video_0")
    save_frames(p, fr, formats=["gif", "joblib", "npz"], fps=8)
    back = joblib.load(p + ".joblib")          # what scripts/eval_owl_vit.py:52 does
    assert back.dtype == np.float32 and np.array_equal(back, fr)
    # identical call as the reference => identical bytes
    joblib.dump(fr, str(tmp_path / "ref.joblib"), compress=("bz2", 3))
    assert open(p + ".joblib", "rb").read() == open(str(tmp_path / "ref.joblib"), "rb").read()
    assert np.array_equal(np.load(p + ".npz")["arr_0"], fr)
    gif = Image.open(p + ".gif")
    assert gif.n_frames == fr.shape[0] and gif.size == (24, 16) and gif.info["duration"] in (120, 125, 130)  # GIF stores centiseconds
    assert gif.info.get("loop") == 0
    with pytest.raises(ValueError, match="Unknown format"):
        save_frames(p, fr, formats="webm")


def test_async_writer_writes_everything_and_reraises(tmp_path):
    import joblib
    from lvd_b200.utils.vis import AsyncFrameWriter
    with AsyncFrameWriter(max_pending=2) as w:
        for i in range(5):
            fr = _frames(i)
            w.submit(str(tmp_path / f"video_{i}"), fr, formats=("joblib",))
            fr[:] = 0  # the writer took its own copy
    for i in range(5):
        assert np.array_equal(joblib.load(str(tmp_path / f"video_{i}.joblib")), _frames(i))
    w2 = AsyncFrameWriter()
    w2.submit(str(tmp_path / "missing_dir" / "video"), _frames(), formats=("npz",))
    with pytest.raises(OSError):
        w2.close()
    assert not os.path.exists(str(tmp_path / "missing_dir"))
