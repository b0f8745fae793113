#!/usr/bin/env python
"""Writes tests/golden/cornell-box.xml: a Mitsuba 0.5 description of the Cornell box with the same numbers as
pathtracer_rs_b200/host/procedural.cpp::build_cornell (themselves the values of the reference's data/cornell-box.xml,
which does not travel with this repo).  The importer test requires import(xml) == build_cornell() bit for bit."""
import os

BSDFS = [("LeftWall", "0.63, 0.065, 0.05"), ("RightWall", "0.14, 0.45, 0.091"), ("Floor", "0.725, 0.71, 0.68"),
         ("Ceiling", "0.725, 0.71, 0.68"), ("BackWall", "0.725, 0.71, 0.68"), ("ShortBox", "0.725, 0.71, 0.68"),
         ("TallBox", "0.725, 0.71, 0.68"), ("Light", "0, 0, 0")]
SHAPES = [
    ("rectangle", "-4.37114e-008 1 4.37114e-008 0 0 -8.74228e-008 2 0 1 4.37114e-008 1.91069e-015 0 0 0 0 1", "Floor", None),
    ("rectangle", "-1 7.64274e-015 -1.74846e-007 0 8.74228e-008 8.74228e-008 -2 2 0 -1 -4.37114e-008 0 0 0 0 1", "Ceiling", None),
    ("rectangle", "1.91069e-015 1 1.31134e-007 0 1 3.82137e-015 -8.74228e-008 1 -4.37114e-008 1.31134e-007 -2 -1 0 0 0 1", "BackWall", None),
    ("rectangle", "4.37114e-008 -1.74846e-007 2 1 1 3.82137e-015 -8.74228e-008 1 3.82137e-015 1 2.18557e-007 0 0 0 0 1", "RightWall", None),
    ("rectangle", "-4.37114e-008 8.74228e-008 -2 -1 1 3.82137e-015 -8.74228e-008 1 0 -1 -4.37114e-008 0 0 0 0 1", "LeftWall", None),
    ("cube", "0.0851643 0.289542 1.31134e-008 0.328631 3.72265e-009 1.26563e-008 -0.3 0.3 -0.284951 0.0865363 5.73206e-016 0.374592 0 0 0 1", "ShortBox", None),
    ("cube", "0.286776 0.098229 -2.29282e-015 -0.335439 -4.36233e-009 1.23382e-008 -0.6 0.6 -0.0997984 0.282266 2.62268e-008 -0.291415 0 0 0 1", "TallBox", None),
    ("rectangle", "0.235 -1.66103e-008 -7.80685e-009 -0.005 -2.05444e-008 3.90343e-009 -0.0893 1.98 2.05444e-008 0.19 8.30516e-009 -0.03 0 0 0 1", "Light", "17, 12, 4"),
]


def main(sunsky=False):
    out = ['<?xml version="1.0" encoding="utf-8"?>', "", '<scene version="0.5.0" >', '\t<integrator type="path" >',
           '\t\t<integer name="maxDepth" value="65" />', "\t</integrator>", '\t<sensor type="perspective" >',
           '\t\t<float name="fov" value="19.5" />', '\t\t<transform name="toWorld" >',
           '\t\t\t<matrix value="-1 0 0 0 0 1 0 1 0 0 -1 6.8 0 0 0 1"/>', "\t\t</transform>", '\t\t<sampler type="sobol" >',
           '\t\t\t<integer name="sampleCount" value="64" />', "\t\t</sampler>", '\t\t<film type="ldrfilm" >',
           '\t\t\t<integer name="width" value="1024" />', '\t\t\t<integer name="height" value="1024" />',
           '\t\t\t<string name="fileFormat" value="png" />', "\t\t</film>", "\t</sensor>"]
    for name, rgb in BSDFS:
        out += [f'\t<bsdf type="twosided" id="{name}" >', '\t\t<bsdf type="diffuse" >', f'\t\t\t<rgb name="reflectance" value="{rgb}"/>',
                "\t\t</bsdf>", "\t</bsdf>"]
    for kind, m, ref, emit in SHAPES:
        out += [f'\t<shape type="{kind}" >', '\t\t<transform name="toWorld" >', f'\t\t\t<matrix value="{m}"/>', "\t\t</transform>",
                f'\t\t<ref id="{ref}" />']
        if emit:
            out += ['\t\t<emitter type="area" >', f'\t\t\t<rgb name="radiance" value="{emit}"/>', "\t\t</emitter>"]
        out += ["\t</shape>"]
    if sunsky:
        out += ['\t<!-- maps to the default environment map (importer/mitsuba.rs:400-418) -->', '\t<emitter type="sunsky" />']
    out += ["</scene>", ""]
    return "\n".join(out)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    open(os.path.join(here, "cornell-box.xml"), "w").write(main(False))
    open(os.path.join(here, "cornell-box-sunsky.xml"), "w").write(main(True))
