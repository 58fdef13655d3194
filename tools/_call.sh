set -x
python tools/time_ewald.py n29_rs1 256 2048 4096 8192
python tools/time_ewald.py n57_rs0.25 64 512 2048
python tools/time_ewald.py n13_rs10 256 4096
