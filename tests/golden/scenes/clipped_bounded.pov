// clipped_by / bounded_by, inverse objects, nested CSG, quadric cone/cylinder forms, orthographic camera
#version 3.7;
global_settings { assumed_gamma 1 }
background { rgb <0.3, 0.4, 0.6> }
camera { orthographic location <0, 4, -10> look_at <0, 1, 0> right x*9.6 up y*5.4 }
light_source { <3, 9, -7> rgb 1 }
plane { y, -0.0078125 pigment { checker rgb 0.85, rgb 0.45 } finish { diffuse 0.7 } }
sphere { <-3, 1, 0>, 1 clipped_by { plane { y, 1.3 } } pigment { rgb <1, 0.4, 0.2> } finish { phong 0.5 } }
sphere { <-3, 1, 0>, 0.9 pigment { rgb <0.2, 0.2, 0.8> } }
quadric { <1, 0, 1>, <0, 0, 0>, <0, 0, 0>, -0.36 clipped_by { box { <-1, 0, -1>, <1, 1.8, 1> } } bounded_by { sphere { <0, 0.9, 0>, 1.2 } }
  pigment { rgb <0.3, 0.8, 0.4> } finish { specular 0.5 } translate <-0.6, 0, 0.5> }
quadric { <1, -1, 1>, <0, 0, 0>, <0, 0, 0>, 0 clipped_by { plane { y, 0 } plane { -y, 1.5 } } pigment { rgb <0.9, 0.8, 0.2> } translate <1.6, 1.5, 0.2> }
union {
  difference { box { <-0.6, 0, -0.6>, <0.6, 1.2, 0.6> } union { sphere { <0, 1.2, 0>, 0.45 } box { <-0.7, 0.3, -0.2>, <0.7, 0.6, 0.2> } } }
  intersection { sphere { <0, 1.7, 0>, 0.5 } plane { y, 1.9 } inverse }
  sphere { <0, 1.7, 0>, 0.2 }
  clipped_by { plane { x, 0.45 } }
  pigment { rgb <0.8, 0.3, 0.7> } finish { phong 0.7 reflection 0.1 }
  translate <3.5, 0, 0.3>
}
