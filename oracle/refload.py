"""Import the reference's OWN Python for the hot path on top of the Faiss shim.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); nothing on the GPU box may call this.  It is used by
oracle/make_golden.py to produce the committed fixtures in tests/golden/ and by
the CPU tests that are skipped when /root/reference is absent.

Three things must be injected before the reference imports:
  1. a ``config`` module -- backend/config.py:46 is a deliberate SyntaxError;
  2. empty stubs for packages that are not installed (skimage, albumentations, flask...);
  3. ``faiss`` -> oracle.faiss_shim.
Nothing is copied: the modules are executed from where they lie.
"""
from __future__ import annotations

import enum
import importlib
import logging
import sys
import types
from pathlib import Path

REFERENCE_BACKEND = Path("/root/reference/backend")


def available() -> bool:
    return (REFERENCE_BACKEND / "kmeans_faiss.py").exists()


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _config_module(models_dir: str = "/tmp/ise_models", n_clusters: int = 200):
    class Method(enum.Enum):
        BOVW = 1
        DNN = 2
        DHASH = 3

    class DnnModels(enum.Enum):
        RESNET = 1
        BiT = 2

    class Config:  # values of backend/config.py:19-109 that the hot path reads
        LOGGING_LEVEL = logging.INFO
        LOGGING_FORMAT = "%(levelname)-5s: @%(funcName)-25s | %(message)s"
        RESIZE_SIZE = 224
        EXTENSIONS = ("*.jpg", "*.jpeg", "*.png")
        NUM_IMAGES_TO_RETURN = 20
        N_JOBS = 1
        DATA_FOLDER_PATH = Path("/tmp/ise_data")
        MODELS_BASE_PATH = Path(models_dir)
        THUMBNAIL_SIZE = 256
        DEVICE = "cpu"
        METHOD = Method.BOVW
        INDEX_TYPE = "l2"
        DHASH_INDEX_PATH = MODELS_BASE_PATH / "dhash_index.pickle"
        DNN_MODEL = DnnModels.RESNET
        DNN_INDEX_PATH = MODELS_BASE_PATH / "resnet50_dnn_index.faiss"
        BOVW_HYPERPARAMETERS_SEARCH = False
        CORNER_DESCRIPTOR = "orb"
        BOVW_CORNER_DESCRIPTIONS_PATH = MODELS_BASE_PATH / "bovw_corner_descriptions.joblib"
        BOVW_KMEANS_INDEX_PATH = MODELS_BASE_PATH / "bovw_kmeans_index.faiss"
        BOVW_PIPELINE_PATH = MODELS_BASE_PATH / "bovw_pipeline.joblib"
        BOVW_INDEX_PATH = MODELS_BASE_PATH / "bovw_index.faiss"
        CLUSTER_EVAL_METHOD = "davies-bouldin"
        CLUSTER_EVAL_SAMPLE_SIZE = 2000
        CLUSTER_EVAL_N_SAMPLES = 10
        NUM_CLUSTERS = n_clusters
        NUM_CLUSTERS_TO_TEST = 3
        MIN_NUM_CLUSTERS = 20
        MAX_NUM_CLUSTERS = 200

    return _stub("config", Config=Config, Method=Method, DnnModels=DnnModels)


_LOADED: dict[str, types.ModuleType] = {}


def load(n_clusters: int = 200):
    """Returns a namespace with the reference modules kmeans_faiss, utils, bag_of_visual_words."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if _LOADED:
        _LOADED["config"].Config.NUM_CLUSTERS = n_clusters
        return types.SimpleNamespace(**_LOADED)
    from oracle import faiss_shim

    sys.modules["faiss"] = faiss_shim
    cfg = _config_module(n_clusters=n_clusters)
    for name in ("skimage", "skimage.feature", "albumentations", "albumentations.pytorch",
                 "flask", "flask_cors"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    sys.modules["skimage.feature"].__dict__.setdefault("hog", None)
    sys.modules["skimage.feature"].__dict__.setdefault("daisy", None)
    sys.modules["albumentations.pytorch"].__dict__.setdefault("ToTensorV2", None)
    sys.path.insert(0, str(REFERENCE_BACKEND))
    try:
        mods = {}
        for name in ("kmeans_faiss", "utils", "descriptors", "bag_of_visual_words"):
            mods[name] = importlib.import_module(name)
    finally:
        sys.path.remove(str(REFERENCE_BACKEND))
    mods["config"] = cfg
    mods["faiss"] = faiss_shim
    _LOADED.update(mods)
    return types.SimpleNamespace(**_LOADED)
