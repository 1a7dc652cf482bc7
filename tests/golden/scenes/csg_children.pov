// CSG over meshes and blobs, bounded_by on CSG children (csg.cpp:128-375 take any ObjectBase child; mesh.cpp:138-262 Inside with
// inside_vector; blob.cpp:596-599 reports every interval of a CSG child)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
background { rgb <0.1, 0.12, 0.2> }
camera { location <0.0137, 3.2071, -8> look_at <0, 0.8, 0> angle 40 right x*16/9 }
light_source { <5, 9, -6> rgb 1 }
light_source { <-6, 4, -3> rgb <0.35, 0.3, 0.25> }
plane { y, -0.0078125 pigment { checker rgb 0.85, rgb 0.3 } finish { ambient 0.1 diffuse 0.7 } }
#declare Octa = mesh {
  triangle { <1,0,0>, <0,1,0>, <0,0,1> }   triangle { <0,1,0>, <-1,0,0>, <0,0,1> }
  triangle { <-1,0,0>, <0,-1,0>, <0,0,1> } triangle { <0,-1,0>, <1,0,0>, <0,0,1> }
  triangle { <0,1,0>, <1,0,0>, <0,0,-1> }  triangle { <-1,0,0>, <0,1,0>, <0,0,-1> }
  triangle { <0,-1,0>, <-1,0,0>, <0,0,-1> } triangle { <1,0,0>, <0,-1,0>, <0,0,-1> }
  inside_vector <0.3, 0.5, 0.8>
}
// mesh as the first child of an intersection, and cut by a sphere
intersection {
  object { Octa scale 1.1 }
  sphere { <0.3, 0.2, -0.2>, 0.95 }
  pigment { rgb <0.9, 0.5, 0.2> } finish { ambient 0.1 diffuse 0.6 phong 0.5 }
  rotate <10, 25, 5> translate <-2.4, 1.1, 0.3>
}
// a box with a mesh-shaped hole (difference = intersection with the inverted mesh)
difference {
  box { <-0.8, -0.8, -0.8>, <0.8, 0.8, 0.8> }
  object { Octa scale 0.95 rotate y*45 }
  pigment { rgb <0.3, 0.7, 0.9> } finish { ambient 0.1 diffuse 0.6 reflection 0.15 }
  rotate <0, -30, 0> translate <0, 0.9, 1.2>
}
// blob children: intersection with a box, and a blob subtracted from a sphere
intersection {
  blob { threshold 0.55 sphere { <-0.5,0,0>, 0.9, 1 } sphere { <0.5,0,0>, 0.9, 1 } cylinder { <0,-0.7,0>, <0,0.7,0>, 0.45, 1 } sphere { <0, 0.3, 0>, 0.5, -0.6 } }
  box { <-1.2, -0.45, -1>, <1.2, 0.5, 1> }
  pigment { rgb <0.4, 0.85, 0.35> } finish { ambient 0.1 diffuse 0.6 specular 0.4 }
  rotate <0, 20, 0> translate <2.5, 0.6, 0.2>
}
difference {
  sphere { 0, 0.8 }
  blob { threshold 0.5 sphere { <0.5,0.3,-0.5>, 0.8, 1 } sphere { <-0.4,0.4,-0.6>, 0.7, 1 } }
  pigment { rgbf <0.95, 0.9, 0.5, 0.6> } finish { ambient 0.05 diffuse 0.3 specular 0.6 roughness 0.02 } interior { ior 1.4 }
  translate <0.4, 0.8, -2.0>
}
// bounded_by on CSG children: a union child and an intersection child each with its own (manual) bounds; one bound is tight
// enough to cut its object off (the reference honours bounded_by literally)
union {
  sphere { <-0.5, 0, 0>, 0.5 bounded_by { box { <-1.05, -0.55, -0.55>, <0.05, 0.55, 0.55> } } }
  intersection { box { <0, -0.4, -0.4>, <0.9, 0.4, 0.4> } sphere { <0.45, 0, 0>, 0.55 } bounded_by { sphere { <0.45, 0, 0>, 0.5 } } }
  cylinder { <-0.2, -0.6, 0>, <-0.2, 0.6, 0>, 0.2 bounded_by { box { <-0.45, -0.3, -0.25>, <0.05, 0.65, 0.25> } } }
  pigment { rgb <0.85, 0.3, 0.6> } finish { ambient 0.1 diffuse 0.6 phong 0.4 }
  rotate <0, -20, 15> translate <-1.0, 0.7, -2.2>
}
