#!/bin/bash
# Dev tool (GPU box): wall time of the unmodified reference CLI and of the CLI bound to libsibgpu on the example genomes.
cd "$(dirname "$0")/../oracle/_ref"
t() { python3 -c "import subprocess,sys,time; s=time.perf_counter(); subprocess.run(sys.argv[1:],stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL); print('%.2f s' % (time.perf_counter()-s))" "$@"; }
for g in Helicobacter_pylori Staphylococcus; do
	for b in Sibelia Sibelia_gpu Sibelia_gpu; do
		rm -rf /tmp/o_$b
		echo "$g $b $(t ./$b -s loose data/$g.fasta -o /tmp/o_$b)"
	done
	cmp /tmp/o_Sibelia/blocks_coords.txt /tmp/o_Sibelia_gpu/blocks_coords.txt && echo "  blocks_coords.txt identical"
done
