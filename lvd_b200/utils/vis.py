"""Mirror of the reference's video writer `utils/vis.py::save_frames` (lines 142-161) -- SURVEY.md section 8(f)
rank 4 -- plus an asynchronous wrapper: the reference writes every video synchronously on the rank that generated it
(`generation/lvd.py:188-196`: joblib bz2-3 + GIF, CPU-bound, ~1 s per 24-frame video) while its GPU idles.

`save_frames(path, frames, formats, fps)` keeps the reference's signature, file names and formats:
  * "joblib": `joblib.dump(frames, f"{path}.joblib", compress=("bz2", 3))` -- the file `scripts/eval_owl_vit.py:52`
    loads with `joblib.load`; identical call, so identical bytes for identical frames;
  * "npz"   : `np.savez_compressed(f"{path}.npz", frames)`;
  * "gif"   : looping GIF at `1000 / fps` ms per frame.  The reference goes through `imageio.mimsave`, whose GIF plugin
    is Pillow; imageio is not installed here, so Pillow is called directly (same encoder, byte equality not claimed);
  * "mp4"   : needs diffusers' `export_to_video` (OpenCV); raises when cv2 is missing rather than silently skipping.
`AsyncFrameWriter` runs the same function on a worker thread so that generation of the next video overlaps the
compression of the previous one; `close()` / the context manager joins and re-raises the first writer error.
"""
import queue
import threading

import numpy as np


def _to_uint8(frames):
    a = np.asarray(frames)
    if a.dtype != np.uint8:  # pipeline output: float32 in [0, 1] (tensor2vid "np")
        a = (np.clip(a, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    return a


def save_frames(path, frames, formats="gif", fps=8):
    if isinstance(formats, (list, tuple)):
        for fmt in formats:
            save_frames(path, frames, fmt, fps)
        return
    if formats == "gif":
        from PIL import Image
        imgs = [Image.fromarray(f) for f in _to_uint8(frames)]
        imgs[0].save(f"{path}.gif", format="GIF", save_all=True, append_images=imgs[1:], loop=0,
                     duration=1000 * 1 / fps)
    elif formats == "mp4":
        try:
            import cv2
        except ImportError as e:  # never a silent skip
            raise RuntimeError("save_frames(..., 'mp4') needs OpenCV (cv2)") from e
        a = _to_uint8(frames)
        h, w = a.shape[1:3]
        vw = cv2.VideoWriter(f"{path}.mp4", cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
        for f in a:
            vw.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
        vw.release()
    elif formats == "npz":
        np.savez_compressed(f"{path}.npz", frames)
    elif formats == "joblib":
        import joblib
        joblib.dump(frames, f"{path}.joblib", compress=("bz2", 3))
    else:
        raise ValueError(f"Unknown format: {formats}")


class AsyncFrameWriter:
    """`submit(path, frames, formats, fps)` returns immediately; one worker thread per writer (one writer per rank)."""

    def __init__(self, max_pending=4):
        self._q = queue.Queue(maxsize=max_pending)
        self._err = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                if self._err is None:
                    save_frames(*item)
            except BaseException as e:  # surfaced by close()
                self._err = e

    def submit(self, path, frames, formats=("gif", "joblib"), fps=8):
        if self._err is not None:
            self.close()
        self._q.put((path, np.array(frames, copy=True), formats, fps))

    def close(self):
        if self._t.is_alive():
            self._q.put(None)
            self._t.join()
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
