rm -f gpurun_out/r2j_ab.log
for o in wf_split=0 wf_split=1; do AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2j_ab.log 2>&1; done
for o in wf_split=0 wf_split=1; do AB_WORLD=8 AB_RANK=3 AB_OPTS=$o AB_TRIS=20000 timeout 300 python scripts/ab_measure.py >> gpurun_out/r2j_ab.log 2>&1; done
cat gpurun_out/r2j_ab.log
timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 400 --timeout-method=thread > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log
