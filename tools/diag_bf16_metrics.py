"""bf16-mode error metrics against the golden fixtures (reference under the bf16 operand shim / fp32)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from clg_vqa_b200 import config as C, synthetic as S, task_utils as TU
from clg_vqa_b200.encoders import BertForVLTasks
for case in ("uc2_L2_B4", "uc2_L12_B8"):
    g = np.load("tests/golden/%s.npz" % case); meta = json.loads(str(g["meta"]))
    cfg = C.uc2_base(meta["layers"], vocab_size=meta["vocab"]); tc = C.gqa_task_cfg()
    model = BertForVLTasks(cfg, tc, ["TASK15"]); model.load_state_dict(S.synth_state_dict(model, seed=0))
    model.set_precision("bf16"); model.to("cuda").eval()
    b = S.synth_batch(meta["B"], seed=meta["batch_seed"], ragged_regions=meta["ragged"], vocab_size=meta["vocab"])
    with torch.no_grad():
        lg = model(b["question"].cuda(), b["features"].cuda(), b["spatials"].cuda(), "TASK15", b["segment_ids"].cuda(), b["input_mask"].cuda(), b["image_mask"].cuda())[0]
        loss, _ = TU.gqa_prior_loss(lg, b["target"].cuda(), b["distances"].cuda(), 10.0)
    lg = lg.double().cpu()
    shim = torch.from_numpy(g["logits_bf16shim"]).double(); f32 = torch.from_numpy(g["logits_fp32"]).double()
    def mx(a, r): return float((a - r).abs().max() / r.abs().max())
    def l2(a, r): return float((a - r).norm() / r.norm())
    print(case, "ours-vs-shim max %.3g l2 %.3g | ours-vs-fp32 max %.3g l2 %.3g | shim-vs-fp32 max %.3g l2 %.3g" % (
        mx(lg, shim), l2(lg, shim), mx(lg, f32), l2(lg, f32), mx(shim, f32), l2(shim, f32)))
    print("   loss ours %.4f shim %.4f fp32 %.4f  rel(ours,shim) %.3g rel(shim,fp32) %.3g" % (
        float(loss), float(g["loss_bf16shim"]), float(g["loss_fp32"]), abs(float(loss) - float(g["loss_bf16shim"])) / float(g["loss_bf16shim"]),
        abs(float(g["loss_bf16shim"]) - float(g["loss_fp32"])) / float(g["loss_fp32"])))
    print("   argmax agree ours/shim %d/%d  shim/fp32 %d/%d" % (int((lg.argmax(1) == shim.argmax(1)).sum()), lg.shape[0], int((shim.argmax(1) == f32.argmax(1)).sum()), lg.shape[0]))
