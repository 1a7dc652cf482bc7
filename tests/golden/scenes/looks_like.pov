// light sources with looks_like geometry.  With a slab tree (this scene) the tree's leaf is the looks_like object itself
// (boundingbox.cpp:337-347): its own flags gate the ray kinds; it never shadows (the parser sets no_shadow on it)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 5 }
camera { location <0, 3, -9> look_at <0, 1, 0> angle 50 }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.2, 0.25, 0.4> } finish { reflection 0.35 diffuse 0.7 } }
light_source { <-3, 4, -1> rgb <1, 0.85, 0.6>
  looks_like { sphere { 0, 0.45 pigment { rgb <1, 0.9, 0.6> } finish { emission 1 diffuse 0 } } } }
light_source { <3, 2.5, 1> rgb <0.4, 0.6, 1>
  looks_like { union { box { -0.3, 0.3 rotate <20, 30, 0> } cone { <0, 0.3, 0>, 0.25, <0, 0.9, 0>, 0 }
               pigment { rgb <0.5, 0.7, 1> } finish { emission 0.8 diffuse 0.2 } bounded_by { sphere { <0, 0.2, 0>, 1.0 } } } } }
// a transparent bulb: the child's interior refracts what lies behind it
light_source { <0, 1.2, -3> rgb 0.5
  looks_like { sphere { 0, 0.5 pigment { rgbf <0.9, 1, 0.9, 0.85> } finish { ambient 0.1 diffuse 0.1 specular 0.6 } interior { ior 1.4 } } } }
// no_reflection on the child: absent from the floor's mirror image (in looks_like_flat.pov, without a tree, it is present)
light_source { <1.2, 0.8, -1.5> rgb 0.3
  looks_like { torus { 0.4, 0.12 rotate 60 * x pigment { rgb <1, 0.3, 0.3> } finish { emission 0.9 } no_reflection } } }
sphere { <-1, 1, 1.5>, 1 pigment { rgb <0.8, 0.3, 0.2> } finish { phong 0.8 reflection 0.2 } }
cylinder { <1.8, 0, 2.5>, <1.8, 2.2, 2.5>, 0.5 pigment { rgb <0.3, 0.7, 0.3> } }
