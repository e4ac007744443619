set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_C2_k20.json 2> gpurun_out/r2f_C2_k20.err
python bench.py --config C2 --steps 200 --warmup 20 > gpurun_out/r2f_C2.json 2> gpurun_out/r2f_C2.err
for c in C1 C3 C4 C5loss; do python bench.py --config $c --steps 100 --warmup 10 > gpurun_out/r2f_$c.json 2> gpurun_out/r2f_$c.err; done
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
bash profiles/capture.sh C2 photo_loss_c3x2 r2f_c2
bash profiles/capture.sh C3 photo_loss_c3x2 r2f_c3
bash profiles/capture.sh C4 photo_loss_nhwc r2f_c4
for t in r2f_c2 r2f_c3 r2f_c4; do ncu -i gpurun_out/${t}_prof.ncu-rep --page raw --csv > gpurun_out/${t}_raw.csv 2>/dev/null; done
python profiles/all_kernels_time.py > gpurun_out/r2f_all_kernels.txt 2>&1
python profiles/feature_time.py > gpurun_out/r2f_feature_time.txt 2>&1
tail -c 600 gpurun_out/r2f_C2_k20.json
