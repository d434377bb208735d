#!/bin/bash
cd /root/repo
W=$(mktemp -d)
python tools/make_main_inputs.py $W --users 6000 > /dev/null
cd $W
export CRX_SHIM_PROFILE=1 CRX_FAKE_SEED=5
for rep in 1 2; do
  s=$(date +%s%N); /root/repo/oracle/_ref/recommendation_crx -d ./tweets.tsv -o ./o1.txt > /dev/null 2> e1.txt; e=$(date +%s%N)
  echo "all GPUs visible: $(( (e - s) / 1000000 )) ms"; grep -E "start-up|k_means|lloyds|VectorReader|create_LSH" e1.txt
done
for rep in 1 2; do
  s=$(date +%s%N); CUDA_VISIBLE_DEVICES=0 /root/repo/oracle/_ref/recommendation_crx -d ./tweets.tsv -o ./o2.txt > /dev/null 2> e2.txt; e=$(date +%s%N)
  echo "CUDA_VISIBLE_DEVICES=0: $(( (e - s) / 1000000 )) ms"; grep -E "start-up" e2.txt
done
nvidia-smi -L | wc -l
