"""Shared pytest plumbing: the `gpu` marker, golden-fixture extraction, oracle import path."""
import json
import lzma
import os
import sys
import tarfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no device exists, whatever -m says."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir(tmp_path_factory):
    """Extract lambda / agona / listeria golden archives (and references/<dataset>.fasta) once per session; returns the
    directory."""
    out = tmp_path_factory.mktemp("golden")
    for name in ("lambda", "agona", "listeria", "references"):
        with tarfile.open(os.path.join(GOLDEN, name + ".tar.xz")) as tar:
            tar.extractall(out, filter="data")
    return str(out)


def load_json_xz(name):
    with lzma.open(os.path.join(GOLDEN, name), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_lines():
    return load_json_xz("ref_lines.json.xz")


@pytest.fixture(scope="session")
def ref_files():
    return load_json_xz("ref_files.json.xz")
