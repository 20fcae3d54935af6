#!/bin/bash
# ncu_export.sh REPORT.ncu-rep : exports the raw and source pages next to the report (gzip) so that a report too large
# for gpurun_out's 64 MiB limit can be analysed on the CPU box; deletes the report when it is larger than $2 MiB (default 30)
rep=$1; lim=${2:-30}; pre=${rep%.ncu-rep}
ncu -i "$rep" --page raw --csv > "$pre.raw.csv" 2>/dev/null
ncu -i "$rep" --page source --csv 2>/dev/null | gzip -9 > "$pre.source.csv.gz"
sz=$(du -m "$rep" | cut -f1)
if [ "$sz" -gt "$lim" ]; then rm -f "$rep"; echo "ncu_export: $rep ($sz MiB) removed after export"; fi
ls -la "$pre".*
