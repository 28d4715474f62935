"""Reader / writer of the reference's stage-I top-K file, the hand-off between stage I and
stage II.  Writer side: src/validate.py:87-94 (Fashion-IQ), :256-263 (CIRR), src/cirr_test_submission.py:123-127.
Reader side: src/data_utils.py:166-179 (Fashion-IQ), :290-305 (CIRR).  The file is a ``torch.save`` dict:

  sorted_index_names  numpy str array [Q, K]   stage-I ranking (reference image removed for CIRR)
  target_names        list[str]  (absent on CIRR test1)      index_names  list[str]
  labels              bool tensor [Q, K]                       split        str
  group_labels        bool tensor [Q, 5]   (CIRR)              dress_types  str  (Fashion-IQ)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


def make_topk_dict(sorted_index_names, index_names: List[str], split: str, target_names: Optional[List[str]] = None,
                   labels=None, group_labels=None, dress_types: Optional[str] = None, k: Optional[int] = None) -> Dict:
    names = np.asarray(sorted_index_names)
    k = names.shape[1] if k is None else k
    d = {"sorted_index_names": names[:, :k], "index_names": list(index_names), "split": split}
    if target_names is not None:
        d["target_names"] = list(target_names)
    if labels is None and target_names is not None:
        labels = torch.from_numpy(names == np.asarray(target_names)[:, None])
    if labels is not None:
        d["labels"] = torch.as_tensor(labels)[:, :k].bool()
    if group_labels is not None:
        d["group_labels"] = torch.as_tensor(group_labels).bool()
    if dress_types is not None:
        d["dress_types"] = dress_types
    return d


def save_topk(path: str, d: Dict) -> None:
    torch.save(d, path)


def load_topk(path: str, K: int, split: str, dress_type: Optional[str] = None, index_names: Optional[List[str]] = None,
              target_names: Optional[List[str]] = None) -> Dict:
    """Same sanity checks as the reference datasets (src/data_utils.py:169-171,293-303): ``dress_type`` given -> the
    Fashion-IQ reader (:166-179), otherwise the CIRR reader (:290-305), which REQUIRES ``group_labels``.  Returns
    ``K_sorted_index_names`` [Q,K], ``K_labels`` numpy bool [Q,K] (None on test1), ``K_group_labels``,
    ``K_target_names``, ``K_index_names``, ``K``."""
    f = torch.load(path, weights_only=False)
    assert K <= f["sorted_index_names"].shape[-1]                       # :169, :293
    assert f["split"] == split                                          # :171, :294
    if dress_type is not None:
        assert f["dress_types"] == dress_type                           # :170
    if index_names is not None:
        assert f["index_names"] == list(index_names), "Something is wrong."          # :297
    out = {"K": K, "K_sorted_index_names": f["sorted_index_names"][:, :K], "K_index_names": f["index_names"],
           "K_labels": None, "K_group_labels": None, "K_target_names": f.get("target_names")}
    if split != "test1":
        out["K_labels"] = f["labels"][:, :K].numpy()                    # :174, :300
        if dress_type is None:                                          # CIRR reader: the key is read unconditionally (:301)
            out["K_group_labels"] = f["group_labels"].numpy()
        if target_names is not None:
            assert out["K_target_names"] == list(target_names), "Something is wrong."  # :303
    return out


def cirr_submission_dicts(pairs_id, K_sorted_index_names, order, group_members, group_order):
    """src/cirr_test_submission_stage2.py:93-106: top-50 re-ranked names and top-3 subset names per pair id.
    ``order`` / ``group_order``: argsort(descending) rows from ``cir_rerank_sort``."""
    names = np.take_along_axis(np.asarray(K_sorted_index_names), np.asarray(order), axis=1)
    gnames = np.take_along_axis(np.asarray(group_members), np.asarray(group_order), axis=1)
    assert gnames.shape[1] == 5
    sub = {"version": "rc2", "metric": "recall"}                        # :50-53
    gsub = {"version": "rc2", "metric": "recall_subset"}                # :54-57
    sub.update({str(int(p)): row[:50].tolist() for p, row in zip(pairs_id, names)})
    gsub.update({str(int(p)): row[:3].tolist() for p, row in zip(pairs_id, gnames)})
    return sub, gsub
