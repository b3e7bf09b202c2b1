"""CPU: the documents cite evidence by path; every cited file under profiles/, scripts/, tests/ and the package must exist."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ("DESIGN.md", "BASELINE.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md"))


def cited_paths(text):
    for tok in re.findall(r"`([^`\s]+)`", text):
        tok = tok.split("::")[0].rstrip(".,;:)")
        if re.match(r"^(profiles|scripts|tests|oracle|include|shim|pipeline_pointcloud_b200)/[\w./\-]+$", tok) and "*" not in tok and "…" not in tok:
            yield tok


def test_cited_files_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        base = os.path.dirname(os.path.join(ROOT, doc))
        for p in cited_paths(text):
            if p.startswith("oracle/_ref") or p.endswith("/"):      # built artefacts / directories
                continue
            if not (os.path.exists(os.path.join(ROOT, p)) or os.path.exists(os.path.join(base, p))):
                missing.append((doc, p))
    assert not missing, missing
