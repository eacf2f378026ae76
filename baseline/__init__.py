"""The reference arm: everything that RUNS THE UNMODIFIED REFERENCE beside libfdms (never part of the product).

    install_ref.py     copies the reference's hot-path modules byte for byte into the git-ignored baseline/_ref/
    refshim.py         makes that tree (or /root/reference in the build container) importable
    ref_runner.py      drives the reference's MultiScaleFaceDetector over the bench workload (both CPU modes, CUDA too)
    backbone_bench.py  a real reference network between the two halves of the hot path: backbone ms, e2e with backbone
    ref_gpu_path.py    the reference's CUDA path against libfdms on the same head maps

Only bench.py's `--impl reference` arm / `cpu_baseline` / `backbone` legs, tests/ and tests/golden/make_golden.py import
from here; nothing under face_detection_multi_scale_b200/, include/ or tools/ does (tests/test_host_logic.py checks)."""
