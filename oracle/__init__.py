"""CPU oracle for the retrieval hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing under
``image_search_engine_b200/`` imports it; the product path fails loudly when
its CUDA extension is missing instead of falling back to anything here.

PARITY UNPINNED: the reference (ManuelZ/image-search-engine) ships no tests,
golden vectors or fixtures for this path and delegates the
arithmetic to the third-party ``faiss-cpu`` package (version unpinned in
``backend/siamese/requirements.txt:2``; not installed here, no network).
``faiss_shim`` therefore restates the *published* Faiss algorithms
(Clustering.cpp, IndexFlat.cpp, utils/distances.cpp, utils/random.cpp,
impl/index_write.cpp, python/extra_wrappers.py, around v1.7.4) and is anchored on the
reference's own call sites.  Everything that is *not* Faiss (np.histogram
binning, OkapiTransformer, chunkIt, create_search_index, FaissKMeans) is
checked against the reference's own Python, imported unmodified from
/root/reference by ``oracle/refload.py`` on top of the shim; the outputs are
committed as fixtures under ``tests/golden/`` by ``oracle/make_golden.py``.
"""
