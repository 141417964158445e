"""The documents only cite what exists: every repository path in DESIGN.md / README.md / INTEGRATION.md resolves to a file
(wildcards: to at least one), and every `file.py::test_name` names a test that is really there."""
import glob
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md"]
TOP = ("tests", "profiles", "tools", "plugin", "oracle", "include", "mediastreamer2_b200", "compat")
# built artefacts and scratch directories are cited too: they need not exist in a fresh checkout
BUILT = ("mediastreamer2_b200/lib/", "plugin/lib/", "oracle/_ref", "oracle/libmsb200oracle.so", "gpurun_out/")


def _cited_paths(text):
    for m in re.finditer(r"`([^`\n]+)`", text):
        tok = m.group(1).strip()
        tok = tok.split("::")[0].split(" ")[0]
        tok = re.sub(r":[0-9][0-9,\- ]*$", "", tok).rstrip(".,;:)")  # file:line citations
        if not tok.startswith(tuple(t + "/" for t in TOP)):
            continue
        if any(tok.startswith(b) for b in BUILT) or "<" in tok or "$" in tok:
            continue
        yield tok


def _expand(tok):
    m = re.search(r"\{([^{}]+)\}", tok)
    if not m:
        return [tok]
    return [x for alt in m.group(1).split(",") for x in _expand(tok[:m.start()] + alt + tok[m.end():])]


@pytest.mark.parametrize("doc", DOCS)
def test_cited_paths_exist(doc):
    text = (ROOT / doc).read_text()
    missing = []
    for tok in sorted(set(_cited_paths(text))):
        for pat in _expand(tok):
            if glob.glob(str(ROOT / pat)) or glob.glob(str(ROOT / (pat + "*"))):
                continue
            # a path of the reference tree (include/mediastreamer2/..., tools/...): checked there when the tree is present
            if pat.startswith(("include/mediastreamer2/", "tools/")) and (not REF.exists() or (REF / pat).exists()):
                continue
            missing.append(pat)
    assert not missing, f"{doc} cites paths that do not exist: {missing}"


@pytest.mark.parametrize("doc", DOCS)
def test_cited_tests_exist(doc):
    text = (ROOT / doc).read_text()
    missing = []
    for m in re.finditer(r"`(?:tests/)?(test_[a-z0-9_]+\.py)::(test_[A-Za-z0-9_]+)", text):
        f = ROOT / "tests" / m.group(1)
        name = m.group(2)
        if not f.exists() or not re.search(rf"def {re.escape(name)}\w*\(", f.read_text()):
            missing.append(f"{m.group(1)}::{name}")
    assert not missing, f"{doc} cites tests that do not exist: {missing}"


def test_sources_cite_existing_tests_and_profiles():
    """comments in the kernels, the plugin, the oracles and the headers name tests and profile records: they must exist"""
    test_names = set()
    for f in (ROOT / "tests").glob("*.py"):
        test_names.update(re.findall(r"def (test_\w+)\(", f.read_text()))
    files = list((ROOT / "mediastreamer2_b200" / "csrc").glob("*.cu")) + list((ROOT / "mediastreamer2_b200" / "csrc").glob("*.h")) + \
        list((ROOT / "plugin").glob("*.[ch]")) + list((ROOT / "oracle").glob("*.[ch]")) + list((ROOT / "include").glob("*.h")) + \
        [ROOT / "bench.py", ROOT / "__graft_entry__.py"]
    missing = []
    for f in files:
        text = f.read_text(errors="replace")
        for name in set(re.findall(r"\b(test_[a-z0-9_]{8,})\b", text)):
            if name.endswith(".py") or (ROOT / "tests" / (name + ".py")).exists():
                continue
            if not any(t == name or t.startswith(name) for t in test_names):
                missing.append(f"{f.name}: {name}")
        for path in set(re.findall(r"\b((?:tests|profiles|tools)/[A-Za-z0-9_./*-]+)", text)):
            path = path.rstrip(".,;:)")
            if not glob.glob(str(ROOT / path)) and not glob.glob(str(ROOT / (path + "*"))):
                missing.append(f"{f.name}: {path}")
    assert not missing, missing
