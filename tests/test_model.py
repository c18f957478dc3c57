"""The parallel formulation used on the device (tests/model_pipeline.py) must
produce the reference's suffix array on every family."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests import families, model_pipeline as mp

KAT = families.kat()


@pytest.mark.parametrize("case", KAT["kat"], ids=lambda c: repr(c["text"])[:24])
def test_model_kat(case):
    t = case["text"].encode("utf-8")
    assert mp.build_sa(t) == case["sa"]


@pytest.mark.parametrize("name,data", families.adversarial(), ids=lambda x: x if isinstance(x, str) else "")
def test_model_adversarial(name, data):
    data = data[:6000]
    assert mp.build_sa(data) == oracle.sais(data).tolist()


@settings(max_examples=300, deadline=None)
@given(st.binary(max_size=120))
def test_model_prop_binary(t):
    assert mp.build_sa(t) == oracle.naive_sa(t).tolist()


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="abc", max_size=80))
def test_model_prop_small_alphabet(s):
    t = s.encode()
    assert mp.build_sa(t) == oracle.naive_sa(t).tolist()


@pytest.mark.parametrize("R", [1, 2, 3, 5, 15])
@pytest.mark.parametrize("name,data", families.adversarial(), ids=lambda x: x if isinstance(x, str) else "")
def test_model_multiround_chain(name, data, R):
    """Model of the cascade steps of k_induce6 (the device uses R = 5): playing R chain
    rounds of a bucket in one partition step gives the same suffix array."""
    data = data[:3000]
    assert mp.build_sa(data, multiround=R) == oracle.sais(data).tolist()


@settings(max_examples=200, deadline=None)
@given(st.text(alphabet="ab", max_size=60), st.integers(1, 6))
def test_model_multiround_prop(s, R):
    t = s.encode()
    assert mp.build_sa(t, multiround=R) == oracle.naive_sa(t).tolist()


@pytest.mark.parametrize("name,data", families.adversarial(), ids=lambda x: x if isinstance(x, str) else "")
def test_model_paircount_induce(name, data):
    """Groundwork for large alphabets (NOTES_ROUND1.md, idea c): with (source, destination)
    pair counts the LMS- and L-part-sourced inductions are one upfront partition each."""
    data = data[:4000]
    assert mp.build_sa(data, paircount=True) == oracle.sais(data).tolist()


@settings(max_examples=300, deadline=None)
@given(st.binary(max_size=100))
def test_model_paircount_prop(t):
    assert mp.build_sa(t, paircount=True) == oracle.naive_sa(t).tolist()


@pytest.mark.parametrize("kc", [1, 2, 3, 16])
@pytest.mark.parametrize("name,data", families.adversarial(), ids=lambda x: x if isinstance(x, str) else "")
def test_model_direct_lms_sort(name, data, kc):
    """Direct LMS-suffix sort by character windows (suffix_b200/csrc/lms_sort.cuh): descending
    feed + stable sorts + truncated-members-are-final give the suffix order without a sentinel."""
    data = data[:1500]
    assert mp.build_sa(data, direct_kc=kc) == oracle.sais(data).tolist()


@settings(max_examples=400, deadline=None)
@given(st.text(alphabet="ab\x00", max_size=60), st.integers(1, 5))
def test_model_direct_lms_sort_prop(s, kc):
    t = s.encode()
    assert mp.build_sa(t, direct_kc=kc) == oracle.naive_sa(t).tolist()


@settings(max_examples=200, deadline=None)
@given(st.binary(max_size=80), st.integers(1, 4))
def test_model_direct_lms_sort_prop_binary(t, kc):
    assert mp.build_sa(t, direct_kc=kc) == oracle.naive_sa(t).tolist()
