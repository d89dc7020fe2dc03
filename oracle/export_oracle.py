"""CPU oracle for the training-snapshot -> inference-weights transformation.  TEST INFRASTRUCTURE ONLY.

Restates `get_source_w` of scripts/export_inference_model.py:18-27 on plain tensors, with the same torch expressions:
the re-parameterised tensors are summed in order and divided by sqrt(k), then every output filter is scaled to unit L2 norm.
Pinned against the reference function itself by tests/test_host.py::test_export_matches_reference_copy_weights (build
container: the reference's training Generator is the source there)."""
import numpy as np
import torch


def merged_filter(ws):
    """ws: list of k tensors [cout, cin/groups, kh, kw] (k = 1: a plain `weight`)."""
    w = ws[0]
    if len(ws) > 1:
        for t in ws[1:]:                                   # :20-22
            w = w + t
        w = w / np.sqrt(len(ws))                           # :23
    return w * (w.square().sum(dim=[1, 2, 3]) + 1e-8).rsqrt().reshape(-1, 1, 1, 1)   # :26
