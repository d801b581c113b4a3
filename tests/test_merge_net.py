"""The generated compare-exchange networks of the kNN kernel (riv-slam_b200/csrc/apd_merge_net.cuh, scripts/gen_merge_net.py):
every network sorts / merges correctly on random inputs with duplicates and empty (all-ones) slots, and the committed header is
what the generator produces."""
import importlib.util
import os

from conftest import ROOT


def _gen():
    spec = importlib.util.spec_from_file_location("gen_merge_net", os.path.join(ROOT, "scripts", "gen_merge_net.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_networks_are_correct():
    g = _gen()
    nets = g.networks()
    g.check(nets)   # asserts inside
    assert set(nets) == {12, 14, 19, 24, 36}     # K + 4 for K = 8, 10, 15, 20, 32
    assert len(nets[24][0]) == 60                 # the figure DESIGN.md quotes: 60 compare-exchanges for the k = 20 list


def test_committed_header_is_current(tmp_path):
    g = _gen()
    committed = open(g.OUT).read()
    g.OUT = str(tmp_path / "apd_merge_net.cuh")
    g.emit(g.networks())
    assert open(g.OUT).read() == committed
