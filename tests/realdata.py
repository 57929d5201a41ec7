"""Locator for the reference's real data files (dict.txt, prefix_dictionary.gob, prob_emit.json).

They are Git-LFS stubs in the reference checkout, so the real-data vectors of
tokenizer_test.go (TestCut, TestBuildDAG, TestCutDag, TestViterbi, TestLoadHMM,
TestBuildPrefixDictFromScratch) can only run where somebody has dropped the real files:
    JIEBA_DATA_DIR=/path/to/dir  (or tests/data/)
Each file must have the sha256 of the upstream blob (kat_vectors.REAL_SHA256); anything else -- a stub,
a different dictionary version -- is refused, and the tests skip with the reason."""
import hashlib
import os

import kat_vectors as kv

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = None


def locate():
    """-> ({name: path}, None) when all three files are present and verified, else (None, reason)."""
    global _cache
    if _cache is not None:
        return _cache
    dirs = [d for d in (os.environ.get("JIEBA_DATA_DIR"), os.path.join(_HERE, "data")) if d]
    reason = "set JIEBA_DATA_DIR to a directory holding dict.txt, prefix_dictionary.gob and prob_emit.json"
    for d in dirs:
        paths, ok = {}, True
        for name, want in kv.REAL_SHA256.items():
            p = os.path.join(d, name)
            if not os.path.isfile(p):
                ok, reason = False, "%s: missing (%s)" % (p, reason)
                break
            h = hashlib.sha256()
            with open(p, "rb") as f:
                for chunk in iter(lambda: f.read(1 << 20), b""):
                    h.update(chunk)
            if h.hexdigest() != want:
                size = os.path.getsize(p)
                what = "a Git-LFS pointer stub" if size < 1024 else "another version"
                ok, reason = False, "%s: sha256 %s... is not the reference's %s... (%s)" % (p, h.hexdigest()[:12], want[:12], what)
                break
            paths[name] = p
        if ok:
            _cache = (paths, None)
            return _cache
    _cache = (None, reason)
    return _cache
