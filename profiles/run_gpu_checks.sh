#!/bin/bash
# The standard GPU bundle of a round, run as ONE gpurun call (1 GPU):
#   gpurun --timeout 1500 -- 'bash profiles/run_gpu_checks.sh <tag> [full]'
# Writes everything under gpurun_out/<tag>_*; the summaries that are judged are
# copied by hand into profiles/ afterwards (gpurun_out/ is scratch).
set -u
TAG=${1:-r1}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1

echo "== pytest -m gpu"
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log

echo "== smoke"
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log

echo "== allhic-b200 optimize on the reference fixture (config 1)"
rm -rf /tmp/fx && mkdir -p /tmp/fx && cp tests/golden/test.ids tests/golden/test.clm /tmp/fx/
( time ./allhic_b200/bin/allhic-b200 optimize /tmp/fx/test.ids /tmp/fx/test.clm ) > $OUT/${TAG}_optimize_fixture.stdout 2> $OUT/${TAG}_optimize_fixture.stderr; echo "optimize rc=$?"
tail -4 $OUT/${TAG}_optimize_fixture.stderr; cp /tmp/fx/test.tour $OUT/${TAG}_fixture.tour 2>/dev/null
( time ./allhic_b200/bin/allhic-b200 optimize /tmp/fx/test.ids /tmp/fx/test.clm --skipGA ) > $OUT/${TAG}_optimize_skipga.stdout 2> $OUT/${TAG}_optimize_skipga.stderr; echo "optimize --skipGA rc=$?"
# --resume (optimize.go:44-51): picks up the .tour the run above left, backs it up to .tour.sav, re-activates its contigs
( ./allhic_b200/bin/allhic-b200 optimize /tmp/fx/test.ids /tmp/fx/test.clm --resume --ngen 200 ) > $OUT/${TAG}_optimize_resume.stdout 2> $OUT/${TAG}_optimize_resume.stderr; echo "optimize --resume rc=$? $(grep -c Success $OUT/${TAG}_optimize_resume.stderr) Success, backup: $(ls /tmp/fx/*.sav 2>/dev/null | wc -l)"

echo "== bench (ours, default flags)"
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/${TAG}_bench.json
echo "== bench (reference arm)"
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 $OUT/${TAG}_bench_ref.json
echo "== bench config2 (N=500, P=1k)"
python bench.py --workload config2 > $OUT/${TAG}_bench_config2.json 2> $OUT/${TAG}_bench_config2.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_config2.json

if [ -n "$FULL" ]; then
  CMD="python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-extras"
  echo "== plain run, then ncu launch list of the same command"
  $CMD > $OUT/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launch.log 2>&1
  echo "launch list rc=$?"
  echo "== ncu --set full of the hot kernels (profiles/prof_target.py: eval_m, GA generation, eval_q, flips)"
  T="python profiles/prof_target.py config3"
  $T > $OUT/${TAG}_prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'eval_m_res_kernel|ga_persist_kernel|eval_q_kernel|eval_q_lookup_kernel|pair_scores_kernel|flip_records_kernel|flip_one_stream_kernel|flip_one_kernel' -s 2 -c 16 -f -o $OUT/${TAG}_prof $T > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -2 $OUT/${TAG}_prof_plain.log
  python profiles/prof_target.py config3 converged > $OUT/${TAG}_prof_converged.log 2>&1; tail -2 $OUT/${TAG}_prof_converged.log
fi
echo "== done"
