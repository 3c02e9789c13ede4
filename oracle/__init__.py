"""CPU oracle for the clr-b200 hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import, link or execute anything in this directory, and there
only as the checker (or as the timed CPU baseline), never as the thing shipped.

PARITY UNPINNED: the arithmetic of the reference's hot path lives in
third-party Julia packages that are not vendored under /root/reference
(NeuralNetworkReachability 0.1, LazySets 7, ReachabilityBase 0.3,
ControllerFormats 0.1-0.2, OrdinaryDiffEq 6; ranges from Project.toml:26-39, no
Manifest) and Julia is absent from this image, so the restatement follows the
published algorithms of those packages from recall and is
anchored on the only in-repo known-answer test that exercises DeepZ without the
Taylor-model plant: the four VerticalCAS verdicts
(models/VerticalCAS/VerticalCAS.jl:311-325).
"""
