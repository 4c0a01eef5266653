for s in ra_1080p ra_4k; do
  b=tests/golden/streams/$s.bin
  ILF_TIMING=2 oracle/_ref/DecoderApp_ilf_b200 -b $b -o /dev/null -d 10 2> gpurun_out/dropin_$s.txt > /dev/null
  grep -c ILFTIME gpurun_out/dropin_$s.txt
  grep "ILFTIME" gpurun_out/dropin_$s.txt | sed -n '3,14p'
done
bash tools/dropin_timing.sh
