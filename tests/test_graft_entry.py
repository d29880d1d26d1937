"""The driver's entry point: build() must succeed from the repo state alone (CPU, no GPU)."""


def test_build_entry_point_runs():
    import __graft_entry__ as g
    g.build()
    assert callable(g.smoke)
