"""TEST INFRASTRUCTURE: restatement of the reference's Blake2b transcript (halo2_proofs/src/transcript.rs:14-20 prefixes,
:199-240 Blake2bWrite, :297-315 Challenge255), used to replay the commit phase of create_proof with real Fiat-Shamir
challenges. hashlib's BLAKE2b takes the same 64-byte digest / 16-byte personalisation parameters as blake2b_simd."""
import hashlib

from oracle import pyref as P

PREFIX_CHALLENGE, PREFIX_POINT, PREFIX_SCALAR = 0, 1, 2


class Blake2bWrite:
    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")  # transcript.rs:175-183
        self.proof = bytearray()

    def common_scalar(self, scalar_int):  # :231-236
        self.state.update(bytes([PREFIX_SCALAR]))
        self.state.update(int(scalar_int).to_bytes(32, "little"))

    def common_point(self, affine_limbs, oracle):  # :217-229: x.to_repr() || y.to_repr() (canonical little-endian)
        pt = P.g1_affine_to_ints(affine_limbs.reshape(1, 8))[0]
        assert pt is not None, "cannot write points at infinity to the transcript"
        x, y = pt
        self.state.update(bytes([PREFIX_POINT]))
        self.state.update(x.to_bytes(32, "little"))
        self.state.update(y.to_bytes(32, "little"))

    def write_point(self, affine_limbs, oracle):  # :199-203
        self.common_point(affine_limbs, oracle)
        self.proof += oracle.g1_to_bytes(affine_limbs)  # compressed encoding, derive/curve.rs:635-646

    def write_scalar(self, scalar_int):  # :204-208: common_scalar, then the canonical 32-byte repr goes to the proof
        self.common_scalar(scalar_int)
        self.proof += int(scalar_int).to_bytes(32, "little")

    def squeeze_challenge_scalar(self):  # :208-215 + Challenge255::new (:297-309): from_bytes_wide of the 64-byte digest
        self.state.update(bytes([PREFIX_CHALLENGE]))
        digest = self.state.copy().digest()
        return int.from_bytes(digest, "little") % P.R_MOD
