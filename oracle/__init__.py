"""ORACLE - TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's transition-matrix path (oracle/classic.py, oracle/contrastive.py), the shim
that imports the unmodified reference in the build container (oracle/ref_shim.py) and the script that generated
tests/golden/*.npz from it (oracle/make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product path (audio_video_textures_b200/)
never does (tests/test_host_cpu.py::test_product_path_has_no_cpu_fallback checks it).
"""
