set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -s --durations=15 ) > gpurun_out/r02k_pytest.log 2>&1; grep -i "through\|passed\|failed\|error\|gaussianHill\|real\|slowest" -A0 gpurun_out/r02k_pytest.log | cut -c1-400 | tail -30
grep -A16 "slowest" gpurun_out/r02k_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
