"""Multi-template matching ("NCC-based OCR"; MatchTool/MatchToolDlg.cpp:718-770): the 36 glyph templates
of the reference's Test Images/M12 read on M12_D_Test (fixtures under tests/golden/images/M12, made by make_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_results_match, get_image

HERE = os.path.dirname(os.path.abspath(__file__))
KNOWN_TEXT = "TESTDATA\n123QWERRTYUIPASDFGHJKL\nZXCVNM823529198765400OO\nQAZ789WSX456EDC123\nRFV987TGB654YHN321"


@pytest.fixture(scope="module")
def ocr_case():
    with open(os.path.join(HERE, "golden", "cases.json")) as f:
        return json.load(f)["ocr_m12"]


def _templates():
    return {ch: get_image("M12/" + ch) for ch in O.OCR_LETTERS}


def test_oracle_reads_m12(ocr_case):
    """known answer: the printed text of the reference's own OCR test image (the full 36-glyph oracle read takes
    ~0.16 s on one core of the GPU box, ~0.4 s here: the CPU figure quoted in profiles/README.md)"""
    assert ocr_case["text"] == KNOWN_TEXT
    # a 6-glyph subset keeps the CPU suite short; the full read is what make_golden.py ran
    sub = {ch: get_image("M12/" + ch) for ch in "TESDA1"}
    text, per = O.ocr_read(get_image(ocr_case["src"]), sub, ocr_case["params"])
    assert text.split("\n")[0] == "TESTDATA"
    for ch in sub:
        assert len(per[ch]) == len(ocr_case["results"][ch])


def test_assemble_matches_oracle_on_random_layouts():
    from fastest_image_pattern_matching_b200 import ocr_assemble
    rng = np.random.default_rng(3)
    for _ in range(300):
        n = int(rng.integers(0, 60))
        lines = int(rng.integers(1, 6))
        cs = [(float(rng.uniform(0, 900)), float(rng.integers(0, lines)) * 40 + float(rng.uniform(-4.5, 4.5))) for _ in range(n)]
        labs = [O.OCR_LETTERS[int(rng.integers(0, 36))] for _ in range(n)]
        assert ocr_assemble(cs, labs) == O.ocr_assemble(list(zip(cs, labs)))


@pytest.mark.gpu
def test_gpu_reads_m12_like_the_oracle(ocr_case):
    from fastest_image_pattern_matching_b200 import GlyphReader
    reader = GlyphReader(_templates(), **ocr_case["params"])
    src = get_image(ocr_case["src"])
    text, per = reader.read(src)
    assert text == KNOWN_TEXT
    for ch in O.OCR_LETTERS:
        assert_results_match(per[ch], ocr_case["results"][ch], ordered=False)
    # concurrent multi-template call == one handle at a time
    for ch, m in zip(reader.letters, reader.matchers):
        assert_results_match(m.match(src), per[ch], 0, 0, 0)


@pytest.mark.gpu
def test_match_multi_edge_cases():
    from fastest_image_pattern_matching_b200 import TemplateMatcher, match_multi
    assert match_multi([], np.zeros((8, 8), np.uint8)) == []
    m = TemplateMatcher()
    # not learned -> empty result like TemplateMatcher::match (src/TemplateMatcher.cpp:99-114), not an error
    assert match_multi([m, m], np.zeros((64, 64), np.uint8)) == [[], []]
