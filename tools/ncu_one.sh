# usage: bash tools/ncu_one.sh <case> <kernel regex> <tag>   -- plain run first, then one ncu --set full capture
CASE=$1; KRE=$2; TAG=$3
python tools/prof_conv.py $CASE 5 || exit 1
ncu --set full --import-source on --clock-control none -k regex:$KRE -c 1 -s 2 -o gpurun_out/$TAG -f python tools/prof_conv.py $CASE 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
