for r in 1 2; do
for lib in prev new; do
  if [ $lib = prev ]; then export SEMIC_B200_LIB=$PWD/semic_b200/lib/libsemic_b200_prev.so; else unset SEMIC_B200_LIB; fi
  echo -n "$lib 2Mi points: "; python bench.py --points 2097152 --steps 20 --warmup 5 --no-extra --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz'])"
done; done
for lib in prev new; do
  if [ $lib = prev ]; then export SEMIC_B200_LIB=$PWD/semic_b200/lib/libsemic_b200_prev.so; else unset SEMIC_B200_LIB; fi
  echo -n "$lib 16Mi points: "; python bench.py --steps 6 --warmup 3 --no-extra --no-e2e --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz'])"
done
