"""Generation -> understanding -> multi-round editing through `MingUniVisionInfer`, the sequence of the reference's
demo (mingunivision/test_infer_unified.py), on the B200-native path.

    python examples/infer_unified.py --model <Ming-UniVision-16B-A3B checkpoint dir> --reference-dir <.../mingunivision>

`--reference-dir` is only a place to find the tokenizer DATA (`tokenizer.json`, `tokenizer_config.json`,
`preprocessor_config.json`) when the checkpoint directory does not hold it — the reference keeps it in its
`mingunivision/` directory.  Chat template, image fetching, token expansion and the CFG masks are this package's own
`processing_bailingmm.BailingMMProcessor` (exact mirror, tests/test_processing_cpu.py).  Needs the real checkpoint — there
is none offline, so this flow is exercised in the tests through injected tiny / stubbed models
(tests/test_llm_gpu.py::test_infer_facade_calls_like_the_reference,
tests/test_host_logic_cpu.py::test_facade_round_trip_with_the_processor)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ming_univision_b200.mingunivisioninfer import MingUniVisionInfer  # noqa: E402


def human(*content):
    return [{"role": "HUMAN", "content": list(content)}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--reference-dir", default="./mingunivision")
    ap.add_argument("--prompt", default="Please generate the corresponding image based on the description. A cute girl.")
    args = ap.parse_args()
    agent = MingUniVisionInfer(args.model, reference_dir=args.reference_dir)
    torch.manual_seed(11)  # the RF sampler draws its noise with torch.randn, as the reference does

    # 1. text -> image (256 continuous visual tokens, rectified-flow head, pixel decoder) -> gen.png
    print(agent.generate(human({"type": "text", "text": args.prompt}), max_new_tokens=512, output_image_prefix="gen"))
    agent.reset_inner_state()

    # 2. image -> text
    print(agent.generate(human({"type": "image", "image": "gen.png"},
                               {"type": "text", "text": "Please describe the picture in detail."}), max_new_tokens=512))
    agent.reset_inner_state()

    # 3. multi-round in-context editing: the KV cache and the three masks persist across calls (PAST_MODE KEEP / DROP)
    edits = ["Change the color of her cloth to red", "Make her smile"]
    for i, edit in enumerate(edits):
        content = [{"type": "text", "text": f"Given the edit instruction: {edit}, please identify the editing region"}]
        if i == 0:
            content.insert(0, {"type": "image", "image": "gen.png"})
        print(agent.generate(human(*content), max_new_tokens=512, for_edit=True, output_image_prefix=f"edit_round_{i}"))
    agent.reset_inner_state()


if __name__ == "__main__":
    main()
