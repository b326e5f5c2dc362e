"""The reference's doc examples (tests/doc_examples.py) on the device."""
import pytest

import aho_corasick_b200 as ab
import doc_examples

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("example", doc_examples.ALL, ids=lambda f: f.__name__)
def test_doc_example(example):
    example(ab)
