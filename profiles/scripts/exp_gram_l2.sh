cd /root/repo
for n in 40000; do
for cfg in "" "AVTEX_GRAM_ST=1" "AVTEX_GRAM_HINT=1" "AVTEX_GRAM_HINT=2" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=1" "AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=2" "AVTEX_GRAM_GROUP=8" "AVTEX_GRAM_GROUP=8 AVTEX_GRAM_ST=1" "AVTEX_GRAM_GROUP=12 AVTEX_GRAM_ST=1" "AVTEX_GRAM_GROUP=24 AVTEX_GRAM_ST=1" "AVTEX_GRAM_GROUP=32 AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=2"; do
env $cfg python profiles/r02_kernels.py gramsym $n
done; done
python profiles/r02_kernels.py gram5 12536
AVTEX_GRAM_ST=1 python profiles/r02_kernels.py gram5 12536
AVTEX_GRAM_ST=1 AVTEX_GRAM_HINT=1 python profiles/r02_kernels.py gram5 12536
