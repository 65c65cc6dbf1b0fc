# geometry sweep of the panel ADJ kernel (G NCW S C HUB), cora_x1024
for g in "2 10 3 1024 32" "2 10 3 1024 64" "2 10 3 1024 16" "3 6 3 640 32" "2 8 3 1024 32" "1 20 3 2048 32" "2 10 4 768 32"; do
  set -- $g; echo "== G=$1 NCW=$2 S=$3 C=$4 HUB=$5"
  SGRACE_PANEL_G=$1 SGRACE_PANEL_NCW=$2 SGRACE_PANEL_S=$3 SGRACE_PANEL_C=$4 SGRACE_PANEL_LONG=$5 timeout 100 python tools/adj_panel_check.py 1024 16 2>&1 | tail -2
done
