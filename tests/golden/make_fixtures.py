"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only).

* tests/golden/cases/{1,2}/*   : byte-for-byte copies of the reference's golden binaries
                                 (reference tests/cases/{1,2}; format in tests/dwbc_test_util.h:15-28)
* tests/golden/tocabi_model.urdf : reference tests/dyros_tocabi.urdf reduced to what a dynamics
                                 model reader consumes (<link><inertial>, <joint> origin/axis/parent/child/limit);
                                 visuals, collisions and MuJoCo tags are dropped.  Numbers are kept as
                                 the literal strings of the source file.
* tests/golden/dof/tocabi_dof{18,25,45}.urdf : the same reduction of three models of the reference's DoF study
                                 (reference tests/dof_test/dyros_tocabi_dof{18,25,45}.urdf, tests/sp_test/dof_comparison.cpp:41-146)
* tests/golden/variants/tocabi_{ub,wo_head,fixed_head}.urdf : the same reduction of the reference's model variants
                                 (reference tests/dyros_tocabi_{ub,wo_head,fixed_head}.urdf; users: tests/sp_test/rot_test.cpp:19,
                                 tests/sp_test/test_mod.cpp:525)
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



def reduce_urdf(src, dst, name):
    root = ET.parse(src).getroot()
    out = ET.Element("robot", {"name": name})
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
    ET.ElementTree(out).write(dst, xml_declaration=True, encoding="utf-8")


reduce_urdf(os.path.join(REF, "tests", "dyros_tocabi.urdf"), os.path.join(HERE, "tocabi_model.urdf"), "tocabi")
os.makedirs(os.path.join(HERE, "dof"), exist_ok=True)
for dof in (18, 25, 45):
    reduce_urdf(os.path.join(REF, "tests", "dof_test", f"dyros_tocabi_dof{dof}.urdf"), os.path.join(HERE, "dof", f"tocabi_dof{dof}.urdf"), f"tocabi_dof{dof}")
os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
for v in ("ub", "wo_head", "fixed_head"):
    reduce_urdf(os.path.join(REF, "tests", f"dyros_tocabi_{v}.urdf"), os.path.join(HERE, "variants", f"tocabi_{v}.urdf"), f"tocabi_{v}")
print("fixtures written under", HERE)
