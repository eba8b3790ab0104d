"""Extract the in-source test inputs of the reference's JPEG tests into data files.

Run in the build container (needs /root/reference); the outputs are committed so
that tests never read /root/reference at run time.

  fuzz_issue10413.bin       decoder.zig "large image with short data" (504 bytes)
  padded_rst_issue28717.jpg decoder.zig "padded rst marker" (base64 literal)
  ref_fixtures/*.jpeg       copies of src/testdata/*.jpeg and iceberg.jpg (test DATA, not source)
"""
import base64
import glob
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

src = open(os.path.join(REF, "src/jpeg/decoder.zig"), encoding="utf-8").read()

# --- 504-byte fuzz input -------------------------------------------------
m = re.search(r'test "large image with short data".*?const input: \[\]const u8 = &\[_\]u8\{(.*?)\};', src, re.S)
data = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1)))
assert len(data) == 504, len(data)
open(os.path.join(HERE, "fuzz_issue10413.bin"), "wb").write(data)

# --- base64 padded-RST image ---------------------------------------------
m = re.search(r'test "padded rst marker".*?const base64EncodedImage =\n(.*?)\n\s*;', src, re.S)
b64 = "".join(line.strip()[2:] for line in m.group(1).splitlines() if line.strip().startswith("\\\\"))
img = base64.b64decode(b64)
assert img[:2] == b"\xff\xd8"
open(os.path.join(HERE, "padded_rst_issue28717.jpg"), "wb").write(img)

# --- fixtures ---------------------------------------------------------------
dst = os.path.join(HERE, "ref_fixtures")
os.makedirs(dst, exist_ok=True)
for f in glob.glob(os.path.join(REF, "src/testdata/*.jpeg")) + [os.path.join(REF, "iceberg.jpg")]:
    shutil.copy(f, dst)
print("ok", len(data), len(img))
