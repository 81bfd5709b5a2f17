"""Pseudo-label writer (row f3 of SURVEY.md section 8): the caller side of full-frame rendering.

The reference turns every rendered view into files synchronously on the training thread
(joint_train_lightning_net.py:246-250, 755-782): normalise + argmax in eager torch, `.detach().cpu().numpy()` of the
float tensors (54 MB of semantics per 640x480 view cross PCIe), a colour look-up, then three `cv2.imwrite` calls while
the GPU idles.  Here

* ucsa_label_epilogue produces the u8 label map (argmax + 1) and the u8 BGR image on the device (1.2 MB per view),
* both go to a ring of pinned host buffers with asynchronous copies on a side stream,
* a small thread pool waits for the copy's event, builds the colour visualisation and PNG-encodes off the critical
  path (cv2 releases the GIL while encoding), so `submit()` returns as soon as the work is enqueued.

File layout and pixel values are those of the reference's predict step: `<root>/<sub>/nerf_image/<index>.png`
(BGR order on disk via cv2), `nerf_label/<index>.png` (single channel, label = argmax + 1) and
`nerf_label_vis/<index>.png` (palette[label]).
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np
import torch

from . import ops


def default_palette(n_labels: int = 41) -> np.ndarray:
    """[n_labels, 3] u8 RGB colours, label 0 black.  Callers that want the reference's look pass
    nr4seg.dataset.ngp_utils.nyu40_colour_code instead (any [>= C + 1, 3] table works)."""
    pal = np.zeros((n_labels, 3), dtype=np.uint8)
    for i in range(1, n_labels):
        h = (i * 0.61803398875) % 1.0  # golden-ratio hue steps: neighbouring labels get distant hues
        s, v = 0.55 + 0.45 * ((i * 7) % 3) / 2.0, 0.65 + 0.35 * ((i * 5) % 4) / 3.0
        k = int(h * 6)
        f = h * 6 - k
        p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
        rgb = [(v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q)][k % 6]
        pal[i] = [int(255 * c) for c in rgb]
    return pal


class _Slot:

    def __init__(self, n_pixels, device):
        self.label = torch.empty(n_pixels, dtype=torch.uint8).pin_memory()
        self.bgr = torch.empty(n_pixels, 3, dtype=torch.uint8).pin_memory()
        self.d_label = torch.empty(n_pixels, dtype=torch.uint8, device=device)
        self.d_bgr = torch.empty(n_pixels, 3, dtype=torch.uint8, device=device)
        self.event = torch.cuda.Event()


class PseudoLabelWriter:

    def __init__(self, root, height, width, palette=None, workers=4, slots=8, device="cuda", write_vis=True):
        import cv2  # noqa: F401  (PNG encoder; fail at construction, not in a worker thread)

        self.root, self.h, self.w = root, int(height), int(width)
        self.device = torch.device(device)
        self.palette = None if not write_vis else np.asarray(default_palette() if palette is None else palette,
                                                             dtype=np.uint8)
        self._free = queue.Queue()
        for _ in range(slots):
            self._free.put(_Slot(self.h * self.w, self.device))
        self._jobs = queue.Queue()
        self._errors = []
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._made_dirs = set()
        self._threads = [threading.Thread(target=self._worker, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()
        self.submitted = 0

    # ------------------------------------------------------------------ producer side (the rendering thread)
    def submit(self, semantics, image, index, subfolder=""):
        """semantics [H*W, C] f32 and image [H*W, 3] f32 of one rendered view (device tensors, e.g. render()'s
        outputs); index = file stem.  Returns immediately; blocks only when all slots are still being written."""
        if semantics.shape[0] != self.h * self.w:
            raise ValueError(f"expected {self.h * self.w} pixels, got {semantics.shape[0]}")
        slot = self._free.get()
        cur = torch.cuda.current_stream(self.device)
        ops.label_epilogue_into(semantics.contiguous(), image.contiguous(), slot.d_label, slot.d_bgr, bgr=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            slot.label.copy_(slot.d_label, non_blocking=True)
            slot.bgr.copy_(slot.d_bgr, non_blocking=True)
            slot.event.record(self._copy_stream)
        self._jobs.put((slot, str(index), subfolder))
        self.submitted += 1

    def close(self):
        """Wait until every submitted view is on disk; re-raises the first worker error."""
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        if self._errors:
            raise self._errors[0]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ consumer side (worker threads)
    def _path(self, subfolder, kind, index):
        d = os.path.join(self.root, subfolder, kind)
        if d not in self._made_dirs:
            os.makedirs(d, exist_ok=True)
            self._made_dirs.add(d)
        return os.path.join(d, index + ".png")

    def _worker(self):
        import cv2

        while True:
            job = self._jobs.get()
            if job is None:
                return
            slot, index, subfolder = job
            try:
                slot.event.synchronize()  # the two copies of this slot have landed
                label = slot.label.numpy().reshape(self.h, self.w)
                bgr = slot.bgr.numpy().reshape(self.h, self.w, 3)
                ok = cv2.imwrite(self._path(subfolder, "nerf_image", index), bgr)
                ok &= cv2.imwrite(self._path(subfolder, "nerf_label", index), label)
                if self.palette is not None:
                    vis = self.palette[label][..., ::-1]  # palette is RGB, cv2 writes BGR
                    ok &= cv2.imwrite(self._path(subfolder, "nerf_label_vis", index), np.ascontiguousarray(vis))
                if not ok:
                    raise OSError(f"cv2.imwrite failed for view {index} under {self.root}")
            except Exception as exc:  # noqa: BLE001 - surfaced by close()
                self._errors.append(exc)
            finally:
                self._free.put(slot)
