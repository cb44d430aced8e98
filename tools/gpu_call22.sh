# round 2, call 22: co-residency A/B again — call 20's chain80s libraries carried a 96-register gather
# (the stray second __launch_bounds__ argument); here the gather is the shipped 64-register one
mkdir -p gpurun_out
run() { # name lib lanes ctas
  name=$1; lib=$2; lanes=$3; ctas=$4
  if [ -n "$lib" ]; then export DRS_LIB_PATH=$PWD/tools/bin/$lib; else unset DRS_LIB_PATH; fi
  if [ -n "$ctas" ]; then export DRS_GATHER_CTAS_PER_SM=$ctas; else unset DRS_GATHER_CTAS_PER_SM; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --lanes $lanes --no_secondary --no_system --no_cpu_baseline --no_e2e --no_sweep > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/r02y_$name.json"))
    print("%-28s value %.2f M  ms/step %.4f  frac %.3f  alone %.4f ms  parity %s  sm_mhz %s" % ("$name", d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_alone"]["ms"], d["parity"]["ok"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$name failed", e)
P
}
run base_l3_c3 "" 3 ""
run chain80s_l3_c3 libdrs_chain80s.so 3 3
run chain80s_l4_c3 libdrs_chain80s.so 4 3
run chain80s_l3_c4 libdrs_chain80s.so 3 4
run chain80s_l2_c3 libdrs_chain80s.so 2 3
run k48u6_chain80s_l4_c4 libdrs_k48u6_chain80s.so 4 4
run base_l4_c3 "" 4 3
run base_again_l3_c3 "" 3 ""
