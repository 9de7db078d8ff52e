"""Stand-in for the CompressAI package, used ONLY to import the reference's own model code in
this container (CompressAI is an un-vendored, un-pinned dependency and is not installed).

TEST INFRASTRUCTURE ONLY.  Nothing here is reference code: the Python layers are loaded BY PATH
from the reference's vendored twins under /root/reference (graphs/layers/gdn_compressai.py,
graphs/layers/entropy_layers_cai.py, utils/bound_ops.py, utils/parametrizers.py) and the two
native functions come from oracle/lbic_oracle.c.
"""


def available_entropy_coders():
    return ["ans"]


def get_entropy_coder():
    return "ans"
