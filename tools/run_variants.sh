for v in trace trace_ldna trace_ldstna trace_ldef; do
  echo "######## $v"
  RVC_B200_LIB=$PWD/rvc_maker_b200/librvc_b200_$v.so timeout 120 python tools/pair_trace.py --quick 2>&1 | cut -c1-420
  RVC_B200_LIB=$PWD/rvc_maker_b200/librvc_b200_$v.so timeout 120 python tools/pair_probe.py 2>&1 | tail -2
done
