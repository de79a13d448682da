#!/bin/bash
# Round profile capture on the GPU box: launch list of the default bench command + one `ncu --set full` capture per hot kernel,
# summarised to text on the box (the .ncu-rep files are too large to travel back together).   bash tools/ncu_round.sh r02
R=${1:-rXX}
O=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${R}_launches.csv $B > /dev/null 2>&1
cap() {  # name, kernel regex, skip, command...
    local name=$1 k=$2 s=$3; shift 3
    ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
    { python tools/ncu_summary.py /tmp/$name.ncu-rep ${TRAFFIC:+--traffic $TRAFFIC}; echo; python tools/ncu_stalls.py /tmp/$name.ncu-rep 12 100000,1000000; echo; python tools/ncu_lines.py /tmp/$name.ncu-rep 30; } > $O/${R}_$name.txt 2>&1
    rm -f /tmp/$name.ncu-rep
}
TRAFFIC=config2_chunk16384_T224 cap crop_cta_config2 bpc_crop_cta 8 $B
cp profiles/crop_traffic.json $O/crop_traffic.json 2>/dev/null
cap crop_prep_config2 bpc_crop_prep 8 $B
cap match_config2 bpc_match_kernel 1 $B
cap match_tri_config2 bpc_match_tri 1 $B
cap crop_cta_300_900 bpc_crop_cta 3 python tools/crop_sweep.py 300 900 224 16384
cap crop_cta_32_96 bpc_crop_cta 3 python tools/crop_sweep.py 32 96 224 16384
cap match_dense_bin bpc_match_kernel 1 python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e --no-crops --dets 200 --scenes 2048
ls -la $O
