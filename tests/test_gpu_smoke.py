"""The driver's smoke entry point must keep working: run it as a test."""
import pytest


@pytest.mark.gpu
def test_graft_entry_smoke():
    import __graft_entry__ as g
    g.smoke()
