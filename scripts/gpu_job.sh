for rep in 1 2; do
echo "== c3 40 chains"; STEPS=10 bash scripts/run_variants.sh 2>&1 | cut -c1-330
done
echo "== c3 5 chains"; STEPS=10 EXTRA="--chains 5" bash scripts/run_variants.sh 2>&1 | cut -c1-330
echo "== c2"; STEPS=10 CFG=c2 bash scripts/run_variants.sh 2>&1 | cut -c1-330
echo "== c4 4 chains"; STEPS=4 CFG=c4 EXTRA="--chains 4" bash scripts/run_variants.sh 2>&1 | cut -c1-330
