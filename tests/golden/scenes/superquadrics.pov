// superellipsoids (superellipsoid.cpp:164-1560): integer and fractional exponents, pinched (e, n > 2) shapes with several hits per ray,
// transformed, transparent with an interior, as operands of difference / intersection (IS_CHILD_OBJECT: every hit is reported)
#version 3.7;
global_settings { assumed_gamma 1.0 max_trace_level 6 }
camera { location <0, 4.5, -10> look_at <0, 0.9, 0> angle 42 }
light_source { <-7, 10, -8> rgb <1, 0.95, 0.9> }
light_source { <8, 5, -5> rgb <0.3, 0.35, 0.5> }
plane { y, 0 pigment { checker rgb <0.85, 0.85, 0.8>, rgb <0.3, 0.35, 0.45> scale 1.5 } finish { reflection 0.1 } }
superellipsoid { <0.25, 0.25> pigment { rgb <0.9, 0.5, 0.2> } finish { phong 0.6 } scale 0.8 translate <-3.8, 0.8, 1.5> }
superellipsoid { <1, 0.4> pigment { rgb <0.3, 0.7, 0.4> } finish { specular 0.4 } scale <0.7, 1, 0.7> rotate <20, 30, 0> translate <-1.9, 1, 1.8> }
superellipsoid { <2.5, 2.5> pigment { rgb <0.5, 0.55, 0.9> } finish { phong 0.5 } scale 1.0 rotate <0, 25, 15> translate <0.2, 1, 2> }
superellipsoid { <0.6, 3.2> pigment { rgb <0.85, 0.3, 0.35> } scale <0.8, 1.1, 0.8> rotate 35 * y translate <2.3, 1.1, 1.6> }
superellipsoid { <2, 0.5> pigment { rgbf <0.9, 1, 0.95, 0.75> } finish { specular 0.5 reflection 0.1 } interior { ior 1.4 }
                 scale 0.9 rotate <10, 40, 0> translate <4.3, 0.9, 1.2> }
difference {
  superellipsoid { <0.3, 0.3> scale 1 }
  superellipsoid { <3, 3> scale 1.15 rotate 45 * y }
  pigment { rgb <0.9, 0.8, 0.3> } finish { specular 0.3 }
  scale 0.9 rotate <0, 20, 0> translate <-2.4, 0.9, -1.8>
}
intersection {
  superellipsoid { <3.5, 0.8> scale 1.1 }
  sphere { 0, 0.85 }
  pigment { rgb <0.4, 0.8, 0.85> } finish { phong 0.5 }
  rotate <15, 10, 0> translate <0.4, 0.9, -2>
}
merge {
  superellipsoid { <0.5, 1.5> scale <0.5, 1, 0.5> }
  superellipsoid { <1.5, 0.5> scale <1, 0.4, 1> }
  pigment { rgbf <1, 0.8, 0.8, 0.6> } interior { ior 1.2 }
  translate <3, 1, -1.8>
}
