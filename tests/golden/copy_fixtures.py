"""Copy the reference's BAM/SAM/definition *test inputs* that the golden cases
read into tests/golden/data/ (run once, in the build container where
/root/reference exists).  These are data fixtures, not reference source code."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from reference_cases import FIXTURES, OWN_GFF_FIXTURES  # noqa: E402

SRC = "/root/reference/tests/data"
DST = os.path.join(HERE, "data")
os.makedirs(DST, exist_ok=True)
for f in FIXTURES:
    shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    print("copied", f)
for name, text in OWN_GFF_FIXTURES.items():  # not reference files: the in-code gene sets of genes.rs's tests, as GFF
    with open(os.path.join(DST, name), "w") as f:
        f.write(text)
    print("wrote", name)
