#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -k "knn" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
for form in dot diff; do
  FZB_KNN_FORM=$form timeout 600 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2o_knn_$form.log 2>&1; echo "knn $form rc=$?"
  tail -4 gpurun_out/r2o_knn_$form.log | cut -c1-260
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2o_knn_launches.csv python tools/bench_knn.py 1000000 32768 20 25 > gpurun_out/r2o_knn_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2o_knn_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:70]:
    print(r[ki][:50], r[gi], r[vi])
PY
