"""Convert an instruction-following json (Alpaca / CoT style: instruction, input, output) into the
``train.json`` / ``test.json`` prompt-response files of the ChatGLM SFT recipe (reference projects/ChatGLM/utils/)."""
import argparse
import json
import os
import random


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--out", default="./data/alpaca")
    ap.add_argument("--test-size", type=int, default=200)
    args = ap.parse_args(argv)
    rows = json.load(open(args.data, encoding="utf-8"))
    random.Random(0).shuffle(rows)
    out = [{"prompt": r.get("instruction", "") + ("\n" + r["input"] if r.get("input") else ""), "response": r.get("output", "")} for r in rows]
    os.makedirs(args.out, exist_ok=True)
    n = min(args.test_size, max(1, len(out) // 10))
    json.dump(out[n:], open(os.path.join(args.out, "train.json"), "w", encoding="utf-8"), ensure_ascii=False)
    json.dump(out[:n], open(os.path.join(args.out, "test.json"), "w", encoding="utf-8"), ensure_ascii=False)
    print(f"train {len(out) - n}, test {n} → {args.out}")


if __name__ == "__main__":
    main()
