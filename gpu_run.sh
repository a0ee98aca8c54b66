cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "allele_counter or flags" 2>&1 | tail -30
