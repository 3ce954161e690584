#!/usr/bin/env python
"""Regenerate tests/golden/ from the reference.  Run in the build container
(needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

What it writes
  vectors/*.xz       the reference's own test vectors (test/square-*.xa and the
                     two source .wav), xz-compressed, byte-identical after
                     decompression (their SHA-1s are pinned in golden.json and
                     equal test/test_decode.sh:24,34,44,54,64,74)
  golden.json        * "reference_tests": the SHA-1s the reference's own test
                       suite asserts (test/test_decode.sh:27-121), re-verified
                       here by running the compiled, unmodified reference
                     * "derived": SHA-1s obtained by running the compiled
                       reference here (PCM-only hashes, encode outputs)
                     * "differential": seeded synthetic streams
                       (bjxa_b200/synth.py) with the SHA-1 of the reference's
                       output for each -- covers what the shipped vectors do not
                       (filter 4, ranges 5/7-15, non-zero header state, long
                       runs, every samples%32, 1-block streams)

Nothing here is read at product run time; tests/ only.
"""
import hashlib
import json
import lzma
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bjxa_b200 import synth  # noqa: E402
from oracle import binding  # noqa: E402

REF_TEST = "/root/reference/test"
OUT = os.path.dirname(os.path.abspath(__file__))

# test/test_decode.sh:24-78 -- (input sha1, decoded-WAV sha1)
REFERENCE_TESTS = {
    "square-stereo-8.xa": ("9fa9edf0ac468129c2e73523df55095a504b8d26",
                           "4b10d39db9abfb75bb3561d7a789ca5afb046c75"),
    "square-mono-8.xa": ("9bdaa12181696bc61a4dfd562edb64a0def2f918",
                         "1c7bdc2f42bd87ebaceb8184312a1857a9f6d8de"),
    "square-stereo-6.xa": ("5241ffdb22617621a6bd7ee9e16055ccb5f59875",
                           "96eac5430bb7a73dc4801449684a4844b9b917c8"),
    "square-mono-6.xa": ("90749ddb703d17d408dd197ff6a877085b80331d",
                         "ce3991eda98db098e45e876944d8324302726a66"),
    "square-stereo-4.xa": ("43e9ddd9afb8208f7bc84cea991fbcd27807a707",
                           "35d8815e712737824c61a02f603145594c0827b7"),
    "square-mono-4.xa": ("02c7ec66ecebda313097462218d9dc05e8886806",
                         "064c48434d77d41c7df3030f3e4a85972dcbac80"),
}
# test/test_decode.sh:121
SATURATION_SHA1 = "56ba3f62bf27ac9fd19cd97bcda06b4db327e612"


def sha1(b: bytes) -> str:
    return hashlib.sha1(b).hexdigest()


def saturation_xa() -> bytes:
    """The hand-written vector of test/test_decode.sh:88-119."""
    hdr = synth.xa_header(66, 32, 44100, 8, 2)
    return hdr + b"\x20" + b"\x7f" * 32 + b"\x20" + b"\x80" * 32


def differential_cases():
    """(kind, params) list; deterministic."""
    cases = []
    key = 1000
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for mix in ("P0", "P1", "P2", "P3"):
                for samples in (1, 31, 32, 33, 1000, 20011):
                    key += 1
                    prev = ((0, 0), (0, 0))
                    if samples in (33, 20011):
                        p = synth.rand_u64(7, key, 4).astype(np.int64)
                        prev = ((int(p[0] % 65536) - 32768, int(p[1] % 65536) - 32768),
                                (int(p[2] % 65536) - 32768, int(p[3] % 65536) - 32768))
                    cases.append({"bits": bits, "channels": ch, "mix": mix,
                                  "samples": samples, "key": key, "prev": prev})
    # every samples % 32
    for r in range(32):
        key += 1
        cases.append({"bits": (4, 6, 8)[r % 3], "channels": 1 + r % 2,
                      "mix": "P2", "samples": 64 + r, "key": key,
                      "prev": ((r, -r), (100 * r, 7))})
    return cases


def main():
    if not binding.have_ref():
        binding.build(ref=True)
    ref = binding.reference_lib()
    orc = binding.Oracle()
    os.makedirs(os.path.join(OUT, "vectors"), exist_ok=True)
    gold = {"reference_tests": {}, "derived": {"pcm_sha1": {}, "encode_sha1": {}},
            "differential": {"seed": 0xB7A, "decode": [], "encode": []}}

    # 1. the reference's own vectors
    for name in sorted(os.listdir(REF_TEST)):
        if not (name.endswith(".xa") or name.endswith(".wav")):
            continue
        raw = open(os.path.join(REF_TEST, name), "rb").read()
        with open(os.path.join(OUT, "vectors", name + ".xz"), "wb") as f:
            f.write(lzma.compress(raw, preset=9 | lzma.PRESET_EXTREME))
        gold.setdefault("vector_sha1", {})[name] = sha1(raw)
        if name.endswith(".xa"):
            in_sha, wav_sha = REFERENCE_TESTS[name]
            assert sha1(raw) == in_sha, name
            wav = ref.xa_to_wav(raw)
            assert sha1(wav) == wav_sha, (name, sha1(wav))
            assert orc.xa_to_wav(raw) == wav, f"oracle differs on {name}"
            gold["reference_tests"][name] = {"input_sha1": in_sha, "wav_sha1": wav_sha}
            gold["derived"]["pcm_sha1"][name] = sha1(wav[44:])
        else:
            for bits in (4, 6, 8):
                xa = ref.wav_to_xa(raw, bits)
                assert orc.wav_to_xa(raw, bits) == xa, f"oracle differs on {name}/{bits}"
                # every profile byte is 0 (libbjxa.c:679)
                gold["derived"]["encode_sha1"][f"{name}:{bits}"] = sha1(xa)

    # 2. the saturation known-answer vector
    sat = saturation_xa()
    assert sha1(ref.xa_to_wav(sat)) == SATURATION_SHA1
    assert orc.xa_to_wav(sat) == ref.xa_to_wav(sat)
    gold["reference_tests"]["saturation"] = {"wav_sha1": SATURATION_SHA1,
                                             "input_sha1": sha1(sat)}

    # 3. seeded differential cases
    seed = gold["differential"]["seed"]
    for c in differential_cases():
        xa = synth.make_xa(seed, c["key"], c["bits"], c["channels"], c["samples"],
                           c["mix"], c["prev"])
        wav = ref.xa_to_wav(xa)
        assert orc.xa_to_wav(xa) == wav, c
        gold["differential"]["decode"].append(dict(c, input_sha1=sha1(xa),
                                                   wav_sha1=sha1(wav)))
    key = 5000
    for bits in (4, 6, 8):
        for ch in (1, 2):
            for frames in (1, 31, 32, 33, 777, 20011):
                key += 1
                pcm = synth.make_pcm(seed, key, ch, frames)
                wav = synth.riff_header(pcm.size * 2, ch) + pcm.tobytes()
                xa = ref.wav_to_xa(wav, bits)
                assert orc.wav_to_xa(wav, bits) == xa, (bits, ch, frames)
                gold["differential"]["encode"].append(
                    {"bits": bits, "channels": ch, "frames": frames, "key": key,
                     "input_sha1": sha1(wav), "xa_sha1": sha1(xa)})

    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(OUT, "golden.json"),
          len(gold["differential"]["decode"]), "decode cases,",
          len(gold["differential"]["encode"]), "encode cases")


if __name__ == "__main__":
    main()
