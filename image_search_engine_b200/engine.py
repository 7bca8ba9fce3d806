"""Drop-in for the query path of backend/engine.py: run_image_query (:46-65).

The Flask plumbing, thumbnail encoding and descriptor extraction (:68-137) are outside the
retrieval core; what is kept is the search call contract:
  * torch tensors are flattened to one (1, d) query (:49-50);
  * ``normalize=True`` L2-normalises the query in place first (:52-53; no reference caller sets it);
  * ``index.search(q, n_images)`` -> flattened distance / id lists (:55-57); inner-product indexes
    return descending scores, L2 indexes ascending squared distances;
  * every id is mapped through ``images_paths`` (:61) -- ids of -1 (k > ntotal)
    are dropped here instead of silently aliasing ``images_paths[-1]``.
The reference reads module globals ``index`` / ``images_paths`` (set under __main__, :110-135); they
can be set on this module the same way or passed explicitly.
"""
from __future__ import annotations

import numpy as np
import torch

from . import faiss_compat as faiss

index = None
images_paths = None
get_image = None  # optional callable path -> base64 thumbnail (utils.py:44-62, UI helper, out of scope)


def run_image_query(image_features, n_images, normalize=False, *, index=None, images_paths=None,
                    get_image=None):
    idx = index if index is not None else globals()["index"]
    paths = images_paths if images_paths is not None else globals()["images_paths"]
    thumb = get_image if get_image is not None else globals()["get_image"]
    if idx is None:
        raise RuntimeError("run_image_query: no index loaded")
    if isinstance(image_features, torch.Tensor) and not image_features.is_cuda:
        image_features = image_features.detach().cpu().numpy().reshape(1, -1)
    elif isinstance(image_features, torch.Tensor):
        image_features = image_features.detach().reshape(1, -1).to(torch.float32)
    else:
        image_features = np.asarray(image_features)
    if normalize:
        faiss.normalize_L2(image_features)
    distances, indices = idx.search(image_features, n_images)
    if isinstance(distances, torch.Tensor):
        distances, indices = distances.cpu().numpy(), indices.cpu().numpy()
    predictions = []
    for dist, i in zip(distances.ravel().tolist(), indices.ravel().tolist()):
        if i < 0:
            continue
        path = paths[i] if paths is not None else i
        predictions.append((dist, thumb(path) if thumb else None, str(path)))
    return predictions


class QueryBatcher:
    """Micro-batching front end for the online query path (engine.py:68-107 serves one upload per request and
    calls ``index.search`` with nq = 1): requests arriving from concurrent server threads within ``max_wait_ms``
    are stacked into ONE ``index.search`` call, so the database is streamed once for the whole batch instead of
    once per request.  ``submit`` returns a ``concurrent.futures.Future`` resolving to the same
    ``[(dist, thumbnail, path), ...]`` list ``run_image_query`` returns; ``query`` is the blocking form.

    Faiss (and this index) evaluates L2 distances with a different formula below nq = 20 (direct sum of squared
    differences) than above (|x|^2 + |y|^2 - 2<x,y>), so near-tied neighbours could swap with the
    batch size.  The batcher therefore pads every batch to at least 20 rows: all requests take the tensor-core
    path, whose per-row result (exactly re-scored, canonically ordered) does not depend on the other rows --
    a request's answer is independent of who else was in its batch (tests/test_gpu_api.py)."""

    def __init__(self, index, images_paths=None, get_image=None, *, max_batch=256, max_wait_ms=2.0):
        import queue
        import threading
        self.index, self.images_paths, self.get_image = index, images_paths, get_image
        self.max_batch, self.max_wait = int(max_batch), float(max_wait_ms) / 1e3
        self._q = queue.Queue()
        self._stop = threading.Event()
        self.batches = []                      # sizes of the batches served (observability / tests)
        self._device = torch.cuda.current_device() if torch.cuda.is_available() else None
        self._t = threading.Thread(target=self._serve, name="ise-query-batcher", daemon=True)
        self._t.start()

    def submit(self, image_features, n_images, normalize=False):
        from concurrent.futures import Future
        if isinstance(image_features, torch.Tensor):
            image_features = image_features.detach().cpu().numpy()
        q = np.ascontiguousarray(np.asarray(image_features, dtype=np.float32).reshape(1, -1))
        if normalize:
            faiss.normalize_L2(q)
        fut = Future()
        self._q.put((q, int(n_images), fut))
        return fut

    def query(self, image_features, n_images, normalize=False):
        return self.submit(image_features, n_images, normalize).result()

    def close(self):
        self._stop.set()
        self._q.put(None)
        self._t.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _serve(self):
        import queue
        import time
        if self._device is not None:
            torch.cuda.set_device(self._device)     # the current device is per thread
        while not self._stop.is_set():
            item = self._q.get()
            if item is None:
                break
            batch = [item]
            deadline = time.monotonic() + self.max_wait
            while len(batch) < self.max_batch:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=left)
                except queue.Empty:
                    break
                if nxt is None:
                    self._stop.set()
                    break
                batch.append(nxt)
            try:
                k = max(n for _, n, _ in batch)
                rows = [q for q, _, _ in batch]
                rows += [rows[-1]] * max(0, faiss.distance_compute_blas_threshold - len(rows))   # one code path
                D, I = self.index.search(np.concatenate(rows, axis=0), k)
                self.batches.append(len(batch))
                for row, (_, n, fut) in enumerate(batch):
                    preds = []
                    for dist, i in zip(D[row, :n].tolist(), I[row, :n].tolist()):
                        if i < 0:
                            continue
                        path = self.images_paths[i] if self.images_paths is not None else i
                        preds.append((dist, self.get_image(path) if self.get_image else None, str(path)))
                    fut.set_result(preds)
            except Exception as exc:      # a failed search fails its requests, not the server thread
                for _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(exc)


def query_index(embedding, index, index_type, n_results):
    """Drop-in for backend/siamese/test_index.py:query_index (:49-71).

    "faiss": L2-normalise the query in place and search the (inner-product) index (:52-56).
    "dict" : the reference's fallback over a pickled [n, d] array -- query divided by its norm, Euclidean
             distance to every row, ascending (:58-69); here the array is searched on the device through an
             L2 index and the square roots are taken on the k winners.
    Returns (indices, distances) like the reference.
    """
    if index_type == "faiss":
        faiss.normalize_L2(embedding)
        distances, indices = index.search(embedding, n_results)
        return indices.ravel().tolist(), distances.ravel().tolist()
    if index_type == "dict":
        q = np.asarray(embedding, dtype=np.float64)
        q = (q / np.linalg.norm(q)).astype(np.float32).reshape(1, -1)
        flat = faiss.IndexFlatL2(q.shape[1])
        flat.add(np.ascontiguousarray(np.asarray(index), dtype=np.float32))
        d2, ids = flat.search(q, n_results)
        keep = ids.ravel() >= 0
        return ids.ravel()[keep], np.sqrt(d2.ravel()[keep])
    raise ValueError(f"unknown index_type {index_type!r}")
