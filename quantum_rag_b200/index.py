"""faiss flat-index files (``IxF2`` / ``IxFI``) straight into a searchable device index.

The reference's ingest tool builds ``faiss.IndexFlatL2(d)``, adds fp32 rows and calls
``faiss.write_index`` (mcp/server/tools/store_in_faiss.py:99-109), and appends the labels to a pickle
(:111-122); nothing in the reference ever searches the result.  This module supplies that missing
retrieval step without faiss: it parses the file format observed in the reference's own fixture
(mcp/piers_morgan_faiss_index.faiss) and hands the rows to the tcgen05 search.

    "IxF2" | int32 d | int64 ntotal | int64 dummy | int64 dummy | uint8 is_trained | int32 metric_type
           | uint64 count (= d * ntotal) | count x fp32 (row major, little endian)        -- 45-byte header
    "IxFI" is the same layout with metric_type 0 (inner product); metric_type 1 is squared L2.
"""
from __future__ import annotations

import io
import json
import os
import pickle
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HEADER = struct.Struct("<4siqqqBiQ")
FAISS_METRIC_IP, FAISS_METRIC_L2 = 0, 1


def parse_ixf(buf: bytes) -> Tuple[np.ndarray, int]:
    """(vectors [ntotal, d] fp32 view of ``buf``, faiss metric_type)."""
    if len(buf) < _HEADER.size:
        raise ValueError("not a flat faiss index: file shorter than its header")
    magic, d, ntotal, _d1, _d2, _trained, metric_type, count = _HEADER.unpack_from(buf, 0)
    if magic not in (b"IxF2", b"IxFI"):
        raise ValueError(f"not a flat faiss index: magic {magic!r}")
    if d <= 0 or ntotal < 0 or count != d * ntotal or len(buf) < _HEADER.size + 4 * count:
        raise ValueError("corrupt flat index: sizes do not match the header")
    x = np.frombuffer(buf, dtype="<f4", count=count, offset=_HEADER.size).reshape(ntotal, d)
    return x, int(metric_type)


def dump_ixf(vectors: np.ndarray, metric_type: int = FAISS_METRIC_L2) -> bytes:
    """The bytes ``faiss.write_index(IndexFlat(d, metric))`` produces for these rows."""
    x = np.ascontiguousarray(vectors, dtype="<f4")
    if x.ndim != 2:
        raise ValueError("vectors must be [ntotal, d]")
    ntotal, d = x.shape
    magic = b"IxF2" if metric_type == FAISS_METRIC_L2 else b"IxFI"
    return _HEADER.pack(magic, d, ntotal, 1 << 20, 1 << 20, 1, metric_type, d * ntotal) + x.tobytes()


class _LabelsOnly(pickle.Unpickler):
    """The metadata side-car is a plain ``list[str]`` (store_in_faiss.py:111-122); refuse anything else."""

    def find_class(self, module, name):          # no globals are needed for list / str / int / dict of those
        raise pickle.UnpicklingError(f"metadata pickle references {module}.{name}; only plain lists are accepted")


def load_metadata(path: str) -> List:
    with open(path, "rb") as fh:
        data = _LabelsOnly(io.BytesIO(fh.read())).load()
    if not isinstance(data, list):
        raise ValueError("metadata side-car is not a list")
    return data


# ------------------------------------------------------------------------------------------------
# JSON side-car: the safe replacement for the pickle the ingest tool appends to (store_in_faiss.py:111-122).
#   {"format": "qrag-labels", "version": 1, "count": N, "labels": [label of row 0, label of row 1, ...]}
# A label is a string (the reference stores "show/episode_sha"), a number, null, or a flat JSON object of those.
# ------------------------------------------------------------------------------------------------
LABELS_FORMAT, LABELS_VERSION = "qrag-labels", 1


def dump_labels(labels: Sequence, path: str) -> None:
    """Write the id -> label side-car (row i of the index carries ``labels[i]``)."""
    doc = {"format": LABELS_FORMAT, "version": LABELS_VERSION, "count": len(labels), "labels": list(labels)}
    tmp = path + ".tmp"
    with open(tmp, "w", encoding="utf-8") as fh:
        json.dump(doc, fh, ensure_ascii=False, allow_nan=False)
    os.replace(tmp, path)                       # the reference rewrites its pickle in place; this never leaves half a file


def load_labels(path: str) -> List:
    with open(path, "r", encoding="utf-8") as fh:
        doc = json.load(fh)
    if not isinstance(doc, dict) or doc.get("format") != LABELS_FORMAT:
        raise ValueError(f"{path}: not a {LABELS_FORMAT} side-car")
    if doc.get("version") != LABELS_VERSION:
        raise ValueError(f"{path}: side-car version {doc.get('version')!r}, this reader understands {LABELS_VERSION}")
    labels = doc.get("labels")
    if not isinstance(labels, list) or doc.get("count") != len(labels):
        raise ValueError(f"{path}: label count does not match the header")
    return labels


def convert_metadata_pickle(pickle_path: str, json_path: Optional[str] = None) -> str:
    """One-shot converter: the reference's ``*_metadata.pkl`` (a plain list, read with the restricted unpickler)
    -> the JSON side-car next to it.  Returns the path written."""
    labels = load_metadata(pickle_path)
    if json_path is None:
        root = pickle_path[:-len("_metadata.pkl")] if pickle_path.endswith("_metadata.pkl") else os.path.splitext(pickle_path)[0]
        json_path = root + "_labels.json"
    dump_labels(labels, json_path)
    return json_path


def sidecar_paths(index_path: str):
    """(json, pickle) side-car paths the ingest convention implies for ``<name>.faiss``."""
    root = index_path[:-len(".faiss")] if index_path.endswith(".faiss") else index_path
    return root + "_labels.json", root + "_metadata.pkl"


class FlatIndex:
    """``faiss.IndexFlat``-like wrapper over the B200 search: ``search(x, k) -> (D, I)``."""

    def __init__(self, vectors, metric_type: int = FAISS_METRIC_L2, labels: Optional[Sequence] = None):
        from . import api
        self.metric_type = int(metric_type)
        self.metric = "l2" if self.metric_type == FAISS_METRIC_L2 else "ip"
        self._tc = api.FlatIndexTC(vectors, self.metric)
        self.d, self.ntotal = self._tc.D, self._tc.N
        self.labels = list(labels) if labels is not None else None
        if self.labels is not None and len(self.labels) != self.ntotal:
            raise ValueError("one label per row expected")

    @classmethod
    def read(cls, path: str, metadata_path: Optional[str] = None) -> "FlatIndex":
        """``metadata_path``: a ``.json`` side-car, or the reference's pickle (restricted unpickler).  Without it the
        JSON side-car next to the index (``<name>_labels.json``) is used when present -- never the pickle implicitly."""
        with open(path, "rb") as fh:
            x, metric_type = parse_ixf(fh.read())
        labels = None
        if metadata_path is None:
            cand = sidecar_paths(path)[0]
            metadata_path = cand if os.path.exists(cand) else None
        if metadata_path is not None:
            labels = load_labels(metadata_path) if metadata_path.endswith(".json") else load_metadata(metadata_path)
        return cls(np.array(x), metric_type, labels)

    def write(self, path: str, with_labels: bool = True) -> None:
        """The ``IxF2`` / ``IxFI`` file faiss would write, plus the JSON side-car when the index has labels."""
        with open(path, "wb") as fh:
            fh.write(dump_ixf(self._tc.X.cpu().numpy(), self.metric_type))
        if with_labels and self.labels is not None:
            dump_labels(self.labels, sidecar_paths(path)[0])

    def search(self, x, k: int):
        """faiss contract: distances (squared L2 ascending / inner product descending) and int64 labels, -1 padded.
        Returned as device tensors (fp64 scores: the exact values the ranking was made on)."""
        if self.ntotal <= 2048 or k > 2048:
            # a corpus of one 2048-row chunk (the reference's fixture has 119 rows): the exact CUDA-core search is two
            # launches against the filter path's eight, and the filter path returns the same bits by construction
            from . import api
            return api.search_topk(x, self._tc.X, k, self.metric)
        return self._tc.search(x, k)

    def search_labels(self, x, k: int):
        """``search`` with ids mapped through the metadata side-car (row label or None for padding)."""
        dist, ids = self.search(x, k)
        if self.labels is None:
            raise ValueError("index has no labels")
        return dist, [[self.labels[i] if i >= 0 else None for i in row] for row in ids.cpu().tolist()]
