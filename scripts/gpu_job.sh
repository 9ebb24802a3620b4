echo "== variants c3 40 chains"; STEPS=12 bash scripts/run_variants.sh 2>&1 | cut -c1-250
echo "== variants c3 40 chains again"; STEPS=12 bash scripts/run_variants.sh 2>&1 | cut -c1-250
echo "== variants c3 5 chains"; STEPS=12 EXTRA="--chains 5" bash scripts/run_variants.sh 2>&1 | cut -c1-250
