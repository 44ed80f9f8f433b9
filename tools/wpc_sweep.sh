# experiment: warps per CTA of k_run_lean (SEMIC_WPC), n_ksub = 24 and 3, 4 Mi points x 73 days and the bench shape
for k in 24 3; do for w in 24 28 32 24 28 32; do echo -n "k=$k WPC=$w: "; SEMIC_WPC=$w python tools/prof_one.py 4194304 73 $k 0 5 | tail -1; done; done
for w in 24 28 32; do echo -n "bench shape k=3 WPC=$w: "; SEMIC_WPC=$w python tools/ab.py 24x0 3 1 6 | tail -1; done
for w in 24 32; do echo -n "bench shape k=24 WPC=$w: "; SEMIC_WPC=$w python tools/ab.py 24x0 24 1 5 | tail -1; done
