# round 2, call 20: can the MLP chain CO-RESIDE with the gather?  (smaller rings: 134 KB instead of
# 166 KB of shared memory, 80 registers; gather capped at 48 registers so that 4 CTAs + the chain fit)
mkdir -p gpurun_out
run() { # name lib lanes ctas
  name=$1; lib=$2; lanes=$3; ctas=$4
  if [ -n "$lib" ]; then export DRS_LIB_PATH=$PWD/tools/bin/$lib; else unset DRS_LIB_PATH; fi
  if [ -n "$ctas" ]; then export DRS_GATHER_CTAS_PER_SM=$ctas; else unset DRS_GATHER_CTAS_PER_SM; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --lanes $lanes --no_secondary --no_system --no_cpu_baseline --no_e2e --no_sweep > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/r02u_$name.json"))
    print("%-28s value %.2f M  ms/step %.4f  frac %.3f  alone %.4f ms  parity %s  sm_mhz %s" % ("$name", d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_alone"]["ms"], d["parity"]["ok"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$name failed", e)
P
}
run base_l3_c3 "" 3 ""
run chain80s_l3_c3 libdrs_chain80s.so 3 3
run chain80s_l3_c4 libdrs_chain80s.so 3 4
run k48u6_chain80s_l3_c4 libdrs_k48u6_chain80s.so 3 4
run k48u8_chain80s_l3_c4 libdrs_k48u8_chain80s.so 3 4
run k48u6_chain80s_l2_c4 libdrs_k48u6_chain80s.so 2 4
run k48u6_chain80s_l3_c5 libdrs_k48u6_chain80s.so 3 5
run k48u6_chain80s_l4_c4 libdrs_k48u6_chain80s.so 4 4
run k48u6_chain80s_l3_c3 libdrs_k48u6_chain80s.so 3 3
run base_again_l3_c3 "" 3 ""
