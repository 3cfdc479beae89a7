#!/bin/bash
# usage: bash tools/sass_summary.sh > profiles/sass_summary.txt
# Static SASS evidence that the hot kernels are tcgen05 / TMA code (cuobjdump -sass of the in-tree sm_100a library):
# UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
# SYNCS = mbarrier ops, REDG = red.global.add.  Template instantiations of one kernel are merged (counts of the first).
SO=domain-transfer-gan_b200/libdtg_b200.so
echo "# $(cuobjdump -lelf $SO 2>/dev/null | grep -c sm_100a) sm_100a cubins in $SO   ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
cuobjdump -sass $SO 2>/dev/null | awk '
/Function : /{name=$3; sub(/^_ZN3dtg[0-9]*/,"",name); sub(/I[LN0-9b_f].*$/,"",name); if(!(name in seen)){seen[name]=1; cur=name} else cur=""}
cur!=""&&/UTCHMMA|UTCQMMA/{a[cur]++} cur!=""&&/UTMALDG/{b[cur]++} cur!=""&&/LDTM/{c[cur]++} cur!=""&&/UTMASTG/{d[cur]++} cur!=""&&/UTCBAR/{e[cur]++} cur!=""&&/SYNCS/{f[cur]++} cur!=""&&/REDG|RED\.E/{g[cur]++}
END{printf "%-34s %8s %8s %6s %8s %7s %6s %5s\n","kernel (dtg::)","UTCHMMA","UTMALDG","LDTM","UTMASTG","UTCBAR","SYNCS","REDG"; for(k in seen) if(a[k]+b[k]+c[k]+d[k]+g[k]>0) printf "%-34s %8d %8d %6d %8d %7d %6d %5d\n",k,a[k],b[k],c[k],d[k],e[k],f[k],g[k]}' | (read h; echo "$h"; sort)
