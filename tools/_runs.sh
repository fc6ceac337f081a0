cd $GRAFT_REPO_ROOT
timeout 600 python tools/gpu_spectral.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_spectral.py -m gpu -x -q 2>&1 | tail -5
