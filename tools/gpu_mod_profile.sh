set -u
OUT=gpurun_out; mkdir -p $OUT
python tools/mod_probe.py 2048 10 > $OUT/r02_mod_plain_mod.log 2>&1; cat $OUT/r02_mod_plain_mod.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mod_shape -s 3 -c 1 -f -o $OUT/r02_mod_prof_mod python tools/mod_probe.py 512 2 > $OUT/r02_mod_ncu_mod.log 2>&1
echo "ncu rc=$?"
python tools/ncu_summary.py $OUT/r02_mod_prof_mod.ncu-rep $OUT/r02_ncu_full_mod_summary.csv
python tools/ncu_src.py $OUT/r02_mod_prof_mod.ncu-rep 40 > $OUT/r02_ncu_src_mod.txt 2>&1
ncu -i $OUT/r02_mod_prof_mod.ncu-rep --page details --csv > $OUT/r02_mod_details.csv 2>&1
rm -f $OUT/r02_mod_prof_mod.ncu-rep
python tools/write_bw_probe.py
