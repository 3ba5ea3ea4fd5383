"""ncu helper: the kernels of the AR visual-token step (B = 3 CFG rows: LLM step on a depth-reduced true-width Bailing-MoE,
vis_head, RF sampler, cached semantic-decoder step, linear_proj) on the eager path, bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --csv python tools/profile_ar_step.py
AR_LAYERS (default 4) LLM layers; the profiled region holds TWO token steps behind a 104-token prefill."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ming_univision_b200 import ops, synthetic  # noqa: E402
from ming_univision_b200.mingtok import MingTokConfig  # noqa: E402
from ming_univision_b200.modeling_bailing_moe import BailingMoeConfig  # noqa: E402
from ming_univision_b200.modeling_bailingmm import MingUniVisionForConditionalGeneration  # noqa: E402

dev = torch.device("cuda:0")
layers = int(os.environ.get("AR_LAYERS", "4"))
llm_cfg = dict(synthetic.LLM_CONFIG, num_hidden_layers=layers, num_image_tokens_for_gen=1)
m = MingUniVisionForConditionalGeneration.on_device(BailingMoeConfig(**llm_cfg), MingTokConfig(**synthetic.MINGTOK_CONFIG),
                                                    synthetic.VISHEAD_CONFIG, dev)
synthetic.init_on_device(m, 0)
ids, um, tm, img_u8 = bench.round_inputs(llm_cfg, 0)
px = ops.image_preprocess(img_u8.to(dev), bench.SIZE, bench.SIZE, out_dtype=torch.bfloat16)
m.model.use_cuda_graph = False
m.model.diffloss.use_cuda_graph = False


def run():
    return m.generate_image_from_prompt(ids.to(dev), pixel_values=px, uncond_attention_mask=um.to(dev),
                                        text_uncond_attention_mask=tm.to(dev))


run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
