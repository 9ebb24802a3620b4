echo "== variants c3 40 chains"; STEPS=6 bash scripts/run_variants.sh 2>&1 | cut -c1-250
echo "== variants c3 5 chains"; STEPS=6 EXTRA="--chains 5" bash scripts/run_variants.sh 2>&1 | cut -c1-250
echo "== variants c2"; CFG=c2 STEPS=6 bash scripts/run_variants.sh 2>&1 | cut -c1-250
echo "== variants c5 4 chains"; CFG=c5 STEPS=3 EXTRA="--chains 4" bash scripts/run_variants.sh 2>&1 | cut -c1-250
