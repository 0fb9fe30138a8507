for pf in 0 16 32 64; do echo "PF=$pf"; OMCHAT_B200_MEGA_PF=$pf timeout 300 python tools/prof_mega.py 28 1 1200 2>&1 | sed -n '1,9p;11p'; done
