"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only).

* tests/golden/cases/{1,2}/*   : byte-for-byte copies of the reference's golden binaries
                                 (reference tests/cases/{1,2}; format in tests/dwbc_test_util.h:15-28)
* tests/golden/tocabi_model.urdf : reference tests/dyros_tocabi.urdf reduced to what a dynamics
                                 model reader consumes (<link><inertial>, <joint> origin/axis/parent/child/limit);
                                 visuals, collisions and MuJoCo tags are dropped.  Numbers are kept as
                                 the literal strings of the source file.
"""
import os
import shutil
import sys
import xml.etree.ElementTree as ET

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

for case in ("1", "2"):
    dst = os.path.join(HERE, "cases", case)
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "tests", "cases", case))):
        shutil.copyfile(os.path.join(REF, "tests", "cases", case, f), os.path.join(dst, f))

root = ET.parse(os.path.join(REF, "tests", "dyros_tocabi.urdf")).getroot()
out = ET.Element("robot", {"name": "tocabi"})
for el in root:
    if el.tag == "link":
        ln = ET.SubElement(out, "link", {"name": el.get("name")})
        inert = el.find("inertial")
        if inert is not None:
            i2 = ET.SubElement(ln, "inertial")
            for ch in inert:
                ET.SubElement(i2, ch.tag, dict(ch.attrib))
    elif el.tag == "joint":
        jn = ET.SubElement(out, "joint", {"name": el.get("name"), "type": el.get("type")})
        for ch in el:
            if ch.tag in ("origin", "parent", "child", "axis", "limit"):
                ET.SubElement(jn, ch.tag, dict(ch.attrib))
ET.indent(out, space=" ")
ET.ElementTree(out).write(os.path.join(HERE, "tocabi_model.urdf"), xml_declaration=True, encoding="utf-8")
print("fixtures written under", HERE)
