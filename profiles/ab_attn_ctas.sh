# how much is one attention CTA slowed by its co-resident twin?  attention class of a 50-step decode at 2 CTAs / SM vs 1 CTA / SM
for i in 1 2; do
echo "== two"; timeout 300 python profiles/step_classes.py fp16 50 2>&1 | tail -1
echo "== one"; SELFTOK_ATTN5_CTAS_PER_SM=1 timeout 300 python profiles/step_classes.py fp16 50 2>&1 | tail -1
done
