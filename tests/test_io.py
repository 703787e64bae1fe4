"""On-disk formats: xvecs.py against files written by the reference's own writers
(tests/golden/io/, produced by tests/golden/make_golden.py from dim_red/data.py:8-21 and
search/support_func.h:194-217) and against values returned by the reference's parsers."""
import os

import numpy as np
import pytest

from gbnns_dim_red_b200 import xvecs

IO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")


@pytest.fixture(scope="module")
def exp():
    return dict(np.load(os.path.join(IO, "expected.npz")))


def _bytes(p):
    with open(p, "rb") as f:
        return f.read()


def test_fvecs_writer_matches_reference_bytes(tmp_path, exp):
    p = tmp_path / "a.fvecs"
    xvecs.write_fvecs(p, exp["small_f"])
    assert _bytes(p) == _bytes(os.path.join(IO, "py_writer.fvecs"))
    assert _bytes(p) == _bytes(os.path.join(IO, "cpp_writer.fvecs"))


def test_ivecs_writer_matches_reference_bytes(tmp_path, exp):
    p = tmp_path / "a.ivecs"
    xvecs.write_ivecs(p, exp["small_i"])
    assert _bytes(p) == _bytes(os.path.join(IO, "py_writer.ivecs"))
    xvecs.write_ivecs(p, exp["small_i"].astype(np.uint32))
    assert _bytes(p) == _bytes(os.path.join(IO, "py_writer.ivecs"))


def test_readers(exp):
    assert np.array_equal(xvecs.read_fvecs(os.path.join(IO, "py_writer.fvecs"), d=5, n=7), exp["small_f"])
    assert np.array_equal(xvecs.read_ivecs(os.path.join(IO, "py_writer.ivecs"), d=6), exp["small_i"].astype(np.uint32))
    assert xvecs.read_fvecs(os.path.join(IO, "py_writer.fvecs"), n=3).shape == (3, 5)
    with pytest.raises(ValueError):  # readXvec prints "file error" and exits (support_func.h:181-188)
        xvecs.read_fvecs(os.path.join(IO, "py_writer.fvecs"), d=4)


def test_edges_roundtrip_variable_degree(tmp_path, exp):
    off, ed = xvecs.read_edges(os.path.join(IO, "cpp_writer_edges.ivecs"))
    assert np.array_equal(off, exp["sub_off"]) and np.array_equal(ed, exp["sub_edges"])
    assert len(set(np.diff(off).tolist())) > 1, "fixture should have variable degree"
    p = tmp_path / "e.ivecs"
    xvecs.write_edges(p, off, ed)
    assert _bytes(p) == _bytes(os.path.join(IO, "cpp_writer_edges.ivecs"))
    off2, _ = xvecs.read_edges(os.path.join(IO, "cpp_writer_edges.ivecs"), n=5)
    assert np.array_equal(off2, exp["sub_off"][:6])


def test_edges_fixed_degree_is_ivecs(tmp_path, exp):
    p = tmp_path / "k.ivecs"
    xvecs.write_ivecs(p, exp["small_i"])
    off, ed = xvecs.read_edges(p)
    assert np.array_equal(off, np.arange(8) * 6) and np.array_equal(ed.reshape(7, 6), exp["small_i"])


def test_empty_and_ragged(tmp_path):
    p = tmp_path / "empty.fvecs"
    open(p, "wb").close()
    assert xvecs.read_fvecs(p, d=4).shape == (0, 4)
    off, ed = xvecs.adjacency_from_lists([[1, 2], [], [0]])
    q = tmp_path / "r.ivecs"
    xvecs.write_edges(q, off, ed)
    off2, ed2 = xvecs.read_edges(q)
    assert off2.tolist() == [0, 2, 2, 3] and ed2.tolist() == [1, 2, 0]


def test_params_parser_matches_reference(exp):
    got = xvecs.read_search_params(os.path.join(IO, "params.txt"), "toy")
    for k, v in zip(exp["param_keys"], exp["param_vals"]):
        assert got.get(str(k), "") == str(v), k
    assert xvecs.int_list(got["efs"]) == exp["efs"].tolist()  # atoi semantics: "x7" -> 0
    assert "efs_sp" not in got  # a value with a space makes a 4-token line, which the parser drops


def test_result_line_format():
    line = xvecs.format_result_line("gd_knn", 0.95, 61, 961, 7.5e-06)
    toks = line.split(" ")
    assert len(toks) == 10 and toks[0] == "graph_type" and toks[2] == "acc"  # draw_results.ipynb cell 8
    assert float(toks[9]) == 7.5e-06
