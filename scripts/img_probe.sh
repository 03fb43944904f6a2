python scripts/img_probe.py 20 2
python scripts/img_probe.py 20 1
CHR3D_SA_CLUSTER=6 CHR3D_SA_WARPS=12 python scripts/img_probe.py 20 1
CHR3D_SA_CLUSTER=3 CHR3D_SA_WARPS=12 python scripts/img_probe.py 40 1
CHR3D_SA_CLUSTER=3 CHR3D_SA_WARPS=10 python scripts/img_probe.py 40 1
CHR3D_SA_CLUSTER=3 CHR3D_SA_WARPS=8 python scripts/img_probe.py 40 1
CHR3D_SA_CLUSTER=2 CHR3D_SA_WARPS=12 python scripts/img_probe.py 40 1
python scripts/img_probe.py 40 1
timeout 900 python -m pytest tests/test_perl_frontend.py -m gpu -x -q 2>&1 | tail -3
