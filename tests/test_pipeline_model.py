"""The array formulation the CUDA kernels implement equals the oracle (CPU only, small)."""
import pytest

from oracle import pyoracle as po
from golden_util import load
from pipeline_model import encode_window


@pytest.mark.parametrize("name,n", [("fuzz_mode0", 30), ("fuzz_mode1", 30), ("fuzz_mode2", 30), ("c1_urls", 25)])
def test_model_matches_oracle(name, n):
    g = load(name)
    docs = [po.make_doc(k, v) for k, v in zip(g["keys"][:n], g["vals"][:n])]
    got = encode_window(docs, strict251=True)
    assert got == g["enc"][:n]
    # incremental: the last 5 docs as a new batch over the window of the first n-5
    assert encode_window(docs, first_new=n - 5, strict251=True) == g["enc"][n - 5:n]
