#!/bin/bash
# Static evidence from the built library: Blackwell-native SASS mnemonics (UTC*MMA = tcgen05.mma, UTMALDG = TMA,
# LDTM = tcgen05.ld, SYNCS = mbarrier), the kernels that hold them, and registers / shared memory / stack per kernel.
#   tools/static_evidence.sh > profiles/rNN_static_sass_evidence.txt        (no GPU needed)
set -e
cd "$(dirname "$0")/.."
SO=gta_graph_tensor_acclelrator_for_general_gnn_b200/libgta_b200.so
[ -f "$SO" ] || python -c "import __graft_entry__ as g; g.build()" >/dev/null
echo "# Static evidence from the built library (cuobjdump on libgta_b200.so, $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)"
echo "# regenerate: tools/static_evidence.sh"
echo
echo "## Blackwell-native instructions in SASS (counts over the whole library)"
cuobjdump -sass $SO 2>/dev/null | grep -oE "\b(UTC[A-Z]*MMA[A-Z.0-9_]*|UTMALDG[A-Z.0-9_]*|UTMASTG[A-Z.0-9_]*|UBLKCP[A-Z.0-9_]*|UTCBAR[A-Z.0-9_]*|LDTM[A-Z.0-9_]*|STTM[A-Z.0-9_]*|UTCATOMSWS[A-Z.0-9_.]*|SYNCS[A-Z.0-9_.]*|HMMA[A-Z.0-9_.]*|LDGSTS[A-Z.0-9_.]*)" | sort | uniq -c | sort -rn
echo
echo "## Which kernels hold UTCHMMA / UTMALDG / LDTM"
cuobjdump -sass $SO 2>/dev/null | awk '/Function :/{fn=$3} /UTCHMMA|UTMALDG|LDTM/{c[fn]++} END{for(f in c) print c[f], f}' | sort -rn | while read n f; do echo "$n $(echo $f | c++filt)"; done
echo
echo "## Resource usage per kernel (cuobjdump -res-usage; CUB kernels omitted): REG, SHARED (static), STACK (spill bytes)"
cuobjdump -res-usage $SO 2>/dev/null | awk '/Function /{fn=$2; sub(/:$/,"",fn)} /REG:/{print fn" "$0}' | while read f rest; do echo "$(echo $f | c++filt | cut -c1-110) | $rest"; done | grep -v "cub::"
echo
echo "## Fused compute + exchange: system-scope peer accesses and release/acquire flags in the AGGREGATION kernels"
echo "## (ld.relaxed.sys on IPC-mapped peer tables = LDG...STRONG.SYS, st.release.sys / ld.acquire.sys on the signal blocks)"
cuobjdump -sass $SO 2>/dev/null | awk '/Function :/{fn=$3} /STRONG\.SYS|MEMBAR\.ALL\.SYS|MEMBAR\.SC\.SYS|\.SYS /{c[fn]++} END{for(f in c) print c[f], f}' | sort -rn | head -12 | while read n f; do echo "$n $(echo $f | c++filt | cut -c1-120)"; done
