"""Extracts the long DWFA known-answer inputs from the reference's own unit test
(src/dwfa/dynamic_wfa.rs:452-468, test_big_early_termination) into a small JSON
fixture.  Run in the build container only (/root/reference does not exist on the
GPU box); the output dwfa_big_early_termination.json is committed."""
import json
import os
import re

SRC = "/root/reference/src/dwfa/dynamic_wfa.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dwfa_big_early_termination.json")

text = open(SRC).read()
c1 = re.search(r'let c1 =\s+"([ACGT]+)";', text).group(1)
seq_23 = re.search(r'let seq_23 = "([ACGT]+)";', text).group(1)
json.dump({
    "source": "src/dwfa/dynamic_wfa.rs:452-468 test_big_early_termination",
    "c1": c1, "seq_23": seq_23,
    "expect_update_ed_max": 2, "expect_final_update_ed": 2, "expect_finalize_ed": 5278,
}, open(OUT, "w"), indent=1)
print(len(c1), len(seq_23))
