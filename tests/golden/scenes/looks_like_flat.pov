// looks_like without a bounding tree (two frame-level objects): the loop over SceneData::objects meets the light source object, whose
// flags gate the ray kinds (trace.cpp:84-95) - the child's no_reflection is never looked at - and whose All_Intersections hands the ray
// to the child (lightsource.cpp:83-95)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 5 }
camera { location <0, 3, -9> look_at <0, 1, 0> angle 50 }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.2, 0.25, 0.4> } finish { reflection 0.35 diffuse 0.7 } }
light_source { <1.2, 1.4, -1.5> rgb 1
  looks_like { torus { 1.0, 0.3 rotate 60 * x pigment { rgbf <1, 0.3, 0.3, 0.5> } finish { emission 0.5 specular 0.4 } interior { ior 1.3 } no_reflection
               bounded_by { sphere { 0, 1.35 } } } } }
