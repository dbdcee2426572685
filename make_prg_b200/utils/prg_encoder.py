"""PRG string -> little-endian uint32 vector (make_prg/utils/prg_encoder.py:22-91): bases A,C,G,T ->
1..4, even markers verbatim, an odd marker verbatim on first sight and +1 when it closes the site."""
import re

import numpy as np

BYTES_PER_INT = 4
ENDIANNESS = "little"


class ConversionError(Exception):
    pass


class EncodeError(Exception):
    pass


def to_bytes(integer):
    return integer.to_bytes(BYTES_PER_INT, ENDIANNESS)


class PrgEncoder:
    encoding = {"A": 1, "C": 2, "G": 3, "T": 4}

    def __init__(self, encoding=None):
        if encoding is not None:
            self.encoding = encoding
        self._site_entry_markers = {}
        self._lut = None

    def _dna_to_int(self, input_char):
        input_char = input_char.upper()
        if input_char not in self.encoding:
            raise ConversionError(f"Char '{input_char}' is not in {self.encoding}")
        return self.encoding[input_char]

    def _encode_unit(self, unit):
        if not unit:
            raise EncodeError("Cannot encode an empty string")
        if all(c.upper() in self.encoding for c in unit):
            if self._lut is None:
                self._lut = np.zeros(256, np.int64)
                for ch, v in self.encoding.items():
                    self._lut[ord(ch.upper())] = v
                    self._lut[ord(ch.lower())] = v
            return self._lut[np.frombuffer(unit.encode(), np.uint8)].tolist()
        if unit.isdigit():
            marker = int(unit)
            if marker % 2 == 0:
                return [marker]
            seen = self._site_entry_markers.get(marker, 0) + 1
            self._site_entry_markers[marker] = seen
            if seen > 2:
                raise ValueError(f"Prg error: odd site marker {marker} found >2 times")
            return [marker if seen == 1 else marker + 1]
        raise EncodeError(f"Unit {unit} contains invalid characters")

    def encode(self, prg):
        out = []
        for unit in prg.split():
            out.extend(self._encode_unit(unit))
        return out

    @staticmethod
    def write(encoding, ostream):
        ostream.write(np.asarray(encoding, dtype="<u4").tobytes())
