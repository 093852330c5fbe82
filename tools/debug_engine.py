import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import orc, to_oracle_cfg
from tinygpt_b200 import engine, models
spec = models.SPECS[sys.argv[1]] if len(sys.argv) > 1 else models.TINY_QWEN2
if spec.max_ctx > 512: spec = spec.with_ctx(256)
w = models.synth_weights(spec, seed=0)
table = models.rope_table(spec)
cfg = to_oracle_cfg(spec)
prompt = torch.randint(0, spec.vocab, (9,), generator=torch.Generator().manual_seed(0))
eng = engine.DecodeEngine(spec, {k: v.cuda() for k, v in w.items()}, table)
for trial in range(2):
    eng.reset_cache()
    first = eng.gen_next_token(prompt.view(1, -1).cuda())
    rest = eng.decode(12)
    print("trial", trial, "tokens", first.view(-1).tolist() + rest.tolist())
eng.reset_cache()
lg = eng.forward(prompt.view(1, -1).cuda())[0, -1].float().cpu()
want = orc.forward(cfg, w, prompt.view(1, -1), orc.KVCache(), table, "bf16")[0, -1]
print("logits err", float((lg - want).abs().max()), "argmax gpu-logits", int(orc.argmax_last(lg.view(1, -1))), "oracle", int(orc.argmax_last(want.view(1, -1))))
toks, _ = orc.generate_greedy(cfg, w, prompt, 13, table)
print("oracle tokens", toks.tolist())
