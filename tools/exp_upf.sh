for d in bf16 fp16; do
  python tools/one_up2.py $d 256 128 64; python tools/one_up2.py $d 256 128 32; python tools/one_up2.py $d 256 256 64; python tools/one_up2.py $d 16 256 128
  for c in 6 8 12; do echo "UP=1 per_sm $c"; UP=1 NBE_UPF_PER_SM=$c python tools/one_up2.py $d 256 128 129;  UP=1 NBE_UPF_PER_SM=$c python tools/one_up2.py $d 256 128 65; done
done
python tools/one_up2.py fp32 256 128 64; UP=1 python tools/one_up2.py fp32 256 128 129
