#!/usr/bin/env python3
"""Extracts, from bincode files the reference itself holds, the facts that pin the leaf encodings of its wire format
(run in the container that has /root/reference; writes tests/golden/bincode_pins.json):

  crates/prover/src/vk_map_dummy.bin = bincode(BTreeMap<[KoalaBear; 8], usize>) (deserialised at crates/prover/src/recursion.rs:57,
  worker/node/full/mod.rs:512).  Its keys are [SP1Field::from_canonical_u32(i); 8] -> i (the construction at recursion.rs:72-75), and
  the bytes hold the words i, i, ... : a KoalaBear element is serialised as its CANONICAL u32, little endian, a fixed-size array has no
  length prefix, usize is a u64, and a map starts with its u64 length.  (Montgomery form would show i * 2^32 mod p = 33554430 for i = 1.)
  crates/prover/src/vk_map.bin (real digests): all words < p and the keys ascend as raw words = BTreeMap order on canonical values.
"""
import hashlib, json, os, struct, sys
import numpy as np
REF = "/root/reference/crates/prover/src"
P = 0x7f000001
out = {"source": "crates/prover/src/vk_map_dummy.bin, vk_map.bin (succinctlabs/sp1 v6.4.0)", "generator": "tools/gen_bincode_pins.py"}
b = open(os.path.join(REF, "vk_map_dummy.bin"), "rb").read()
n = struct.unpack("<Q", b[:8])[0]
assert len(b) == 8 + 40 * n
rec = np.frombuffer(b[8:], np.uint8).reshape(n, 40)
keys = rec[:, :32].copy().view("<u4"); idx = rec[:, 32:].copy().view("<u8")[:, 0]
assert (keys == np.arange(n, dtype=np.uint32)[:, None]).all() and (idx == np.arange(n)).all()
out["dummy"] = {"entries": int(n), "bytes": len(b), "sha256": hashlib.sha256(b).hexdigest(), "head_hex": b[:8 + 3 * 40].hex(),
                "statement": "key i = [from_canonical_u32(i); 8] is stored as eight u32 words equal to i, value i as one u64"}
b = open(os.path.join(REF, "vk_map.bin"), "rb").read()
n = struct.unpack("<Q", b[:8])[0]
assert len(b) == 8 + 40 * n
rec = np.frombuffer(b[8:], np.uint8).reshape(n, 40)
keys = rec[:, :32].copy().view("<u4")
t = [tuple(int(x) for x in r) for r in keys]
assert keys.max() < P and all(t[i] < t[i + 1] for i in range(n - 1))
out["vk_map"] = {"entries": int(n), "bytes": len(b), "sha256": hashlib.sha256(b).hexdigest(), "head_hex": b[:8 + 2 * 40].hex(),
                 "max_word": int(keys.max()), "statement": "every key word < p; keys strictly ascending as raw words"}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bincode_pins.json"), "w"), indent=1)
print("ok")
