for wl in jobshop20 toc_dbr simple_det; do
  RU=2000; [ $wl = toc_dbr ] && RU=4001; [ $wl = simple_det ] && RU=1000
  RU=$RU bash tools/ab_bench.sh r02w_$wl "base|MSIM_LIB=manusim_b200/libmsim_base.so|--workload $wl" "bufA+pohi||--workload $wl" "nobufa|MSIM_LIB=manusim_b200/libmsim_nobufa.so|--workload $wl" "base2|MSIM_LIB=manusim_b200/libmsim_base.so|--workload $wl"
done
