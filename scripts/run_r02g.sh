set -x
mkdir -p gpurun_out
timeout 600 python bench.py --config loh1-h100 > gpurun_out/r02g_loh1_h100_n1.json 2> gpurun_out/r02g_loh1_h100_n1.err; cut -c1-900 gpurun_out/r02g_loh1_h100_n1.json; tail -n 3 gpurun_out/r02g_loh1_h100_n1.err
timeout 900 python bench.py --config loh1-h50 > gpurun_out/r02g_loh1_h50_n1.json 2> gpurun_out/r02g_loh1_h50_n1.err; cut -c1-900 gpurun_out/r02g_loh1_h50_n1.json; tail -n 3 gpurun_out/r02g_loh1_h50_n1.err
