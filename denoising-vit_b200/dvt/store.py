"""The on-disk feature store between stage 1 and stage 2.

Format (reference main_img_denoising.py:131-146, read back by dvt/dataset/paired_list_dataset.py:27-43): NPY v1,
float32, C order; `raw_features/<model>/<rel>.npy` holds (h, w, C), `denoised_features/<model>/<rel>.npy` holds
(1, h, w, C); paths follow `dvt.utils.misc.feature_paths`, and an image counts as done when both files exist
(`check_if_file_exists`, dvt/utils/misc.py:325-337).

`FeatureStoreWriter` takes the writes off the per-image critical path: the device-to-host copy goes onto a side stream
into pinned staging buffers and a worker thread does the file I/O, so neither blocks the stage-1 loop (the reference
does `np.save(...cpu().numpy())` in line).  Files are written to a temporary name and renamed, so a killed run never
leaves a half-written file that the resume rule would take for a finished image."""
from __future__ import annotations

import os
import queue
import threading
from typing import List, Optional, Tuple

import numpy as np
import torch


def save_npy_atomic(path: str, array: np.ndarray) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.tmp.{os.getpid()}.{threading.get_ident()}"
    with open(tmp, "wb") as fh:
        np.save(fh, np.ascontiguousarray(array, dtype=np.float32))
    os.replace(tmp, path)


def load_pair(denoised_path: str) -> Tuple[np.ndarray, np.ndarray]:
    """(raw (h, w, C), denoised (h, w, C)) the way the stage-2 dataset reads them (paired_list_dataset.py:30-36)."""
    raw_path = denoised_path.replace("denoised_features", "raw_features")
    return np.load(raw_path).squeeze(), np.load(denoised_path).squeeze()


class FeatureStoreWriter:
    def __init__(self, max_pending: int = 4):
        self._q: "queue.Queue" = queue.Queue()
        self._free: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self._max_pending = max_pending
        self._allocated = 0
        self._slots = threading.Semaphore(max_pending)
        self._lock = threading.Lock()
        self._error: Optional[BaseException] = None
        self._stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._thread = threading.Thread(target=self._work, name="dvt-feature-store", daemon=True)
        self._thread.start()
        self.written: List[Tuple[str, str]] = []

    # ---- producer side ----------------------------------------------------------------------------------------
    def _buffers(self, raw: torch.Tensor, den: torch.Tensor):
        with self._lock:
            for k, (a, b) in enumerate(self._free):
                if a.shape == raw.shape and b.shape == den.shape:
                    return self._free.pop(k)
        pin = torch.cuda.is_available()
        mk = lambda t: torch.empty(t.shape, dtype=torch.float32, pin_memory=pin)  # noqa: E731
        return mk(raw), mk(den)

    def submit(self, raw_path: str, denoised_path: str, raw: torch.Tensor, denoised: torch.Tensor) -> None:
        """raw (h, w, C) and denoised (1, h, w, C), device or host tensors.  Returns as soon as the copies are enqueued."""
        self.raise_if_failed()
        self._slots.acquire()                       # bounds pinned memory when the disk is slower than the GPU
        hr, hd = self._buffers(raw, denoised)
        event = None
        if raw.is_cuda:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                hr.copy_(raw.float(), non_blocking=True)
                hd.copy_(denoised.float(), non_blocking=True)
                event = torch.cuda.Event()
                event.record(self._stream)
            raw.record_stream(self._stream)
            denoised.record_stream(self._stream)
        else:
            hr.copy_(raw.float())
            hd.copy_(denoised.float())
        self._q.put((raw_path, denoised_path, hr, hd, event))

    # ---- worker ------------------------------------------------------------------------------------------------
    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            raw_path, den_path, hr, hd, event = item
            try:
                if event is not None:
                    event.synchronize()
                save_npy_atomic(raw_path, hr.numpy())
                save_npy_atomic(den_path, hd.numpy())
                self.written.append((raw_path, den_path))
            except BaseException as e:  # surfaced by the next submit() / close()
                self._error = e
            finally:
                with self._lock:
                    self._free.append((hr, hd))
                self._slots.release()
                self._q.task_done()

    def raise_if_failed(self):
        if self._error is not None:
            e, self._error = self._error, None
            raise RuntimeError(f"feature store write failed: {e!r}") from e

    def flush(self):
        self._q.join()
        self.raise_if_failed()

    def close(self):
        self.flush()
        self._q.put(None)
        self._thread.join()
