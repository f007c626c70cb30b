"""Extracts the constant tables the reference embeds for the pairing path into a JSON fixture
(tests/golden/reference_pairing_constants.json). Run in the build container, where
/root/reference is mounted; the fixture travels to the GPU box, the reference does not.

  bn256:     src/circuit/bn256_constants.rs  (32-byte little-endian canonical values)
  bls12_381: src/circuit/bls12_381_pairing_chip.rs:58-107 (Fq::from_raw_unchecked = Montgomery
             limbs, R = 2^384; converted to canonical here)
"""
import json
import os
import re

REF = "/root/reference/src/circuit"
P_BLS = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def parse_bn():
    src = open(os.path.join(REF, "bn256_constants.rs")).read()
    out = {}
    out["BN_X"] = int(re.search(r"BN_X: u64 = (\d+);", src).group(1))
    naf = re.search(r"SIX_U_PLUS_2_NAF: \[i8; 65\] = \[(.*?)\];", src, re.S).group(1)
    out["SIX_U_PLUS_2_NAF"] = [int(x) for x in naf.replace("\n", " ").split(",") if x.strip()]
    for name in ["XI_TO_Q_MINUS_1_OVER_2", "FROBENIUS_COEFF_FQ2_C1", "FROBENIUS_COEFF_FQ6_C1", "FROBENIUS_COEFF_FQ6_C2",
                 "FROBENIUS_COEFF_FQ12_C1"]:
        body = re.search(r"pub const " + name + r": [^=]*= (\[.*?\]);\n", src, re.S).group(1)
        nums = [int(x) for x in re.findall(r"\d+", body)]
        assert len(nums) % 32 == 0
        out[name] = [int.from_bytes(bytes(nums[i:i + 32]), "little") for i in range(0, len(nums), 32)]
    return out


def parse_bls():
    src = open(os.path.join(REF, "bls12_381_pairing_chip.rs")).read()
    vals = []
    for m in re.finditer(r"from_raw_unchecked\(\[(.*?)\]\)", src, re.S):
        limbs = [int(x.replace("_", ""), 16) for x in re.findall(r"0x[0-9a-f_]+", m.group(1))]
        assert len(limbs) == 6
        mont = sum(l << (64 * i) for i, l in enumerate(limbs))
        vals.append(mont * pow(1 << 384, -1, P_BLS) % P_BLS)
    x = int(re.search(r"BLS_X: u64 = (0x[0-9a-f_]+);", src).group(1).replace("_", ""), 16)
    # order in the file: fq6 c1 coeff (.c1), fq6 c2 coeff (.c0), fq12 coeff (.c0, .c1)
    return {"BLS_X": x, "FQ6_C1_c1": vals[0], "FQ6_C2_c0": vals[1], "FQ12_C1": [vals[2], vals[3]]}


if __name__ == "__main__":
    out = {"bn256": parse_bn(), "bls12_381": parse_bls()}
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_pairing_constants.json"), "w") as f:
        json.dump(out, f, indent=1)
    print({k: (len(v) if isinstance(v, list) else v) for k, v in out["bn256"].items()}, out["bls12_381"]["BLS_X"])
