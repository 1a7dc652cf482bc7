#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
background { rgb <0.1, 0.12, 0.2> }
camera { perspective location <0, 4, -11> direction <0, 0, 1.6> up <0, 1, 0> right <1.7777777777777777, 0, 0> look_at <0, 1.2, 0> }
light_source { <12, 18, -14> rgb <1, 1, 1> }
light_source { <-10, 8, -6> rgb <0.35, 0.35, 0.45> }
plane { y, -0.0078125 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.2, 0.25, 0.3> } finish { ambient 0.1 diffuse 0.7 } }
// spheres only, bounding hierarchy (default)
blob { threshold 0.6
  sphere { <-3.2, 1.0, 0.0>, 1.3, 1.0 } sphere { <-2.3, 1.6, 0.3>, 1.1, 1.0 } sphere { <-3.8, 1.9, -0.4>, 0.9, 0.8 }
  sphere { <-2.9, 2.4, 0.2>, 0.8, 1.0 } sphere { <-3.0, 1.5, -0.9>, 0.7, -0.6 }
  pigment { rgb <0.9, 0.35, 0.25> } finish { ambient 0.1 diffuse 0.6 phong 0.6 phong_size 50 } }
// cylinders (cylinder + two hemispheres each) and a scaled sphere (ellipsoid), no hierarchy, sturm
blob { threshold 0.5
  cylinder { <0.0, 0.3, 0.0>, <0.0, 2.6, 0.0>, 0.9, 1.0 } cylinder { <-1.0, 1.4, 0.0>, <1.0, 1.4, 0.2>, 0.7, 1.0 }
  sphere { <0.0, 2.9, 0.0>, 1.2, 0.9 scale <1.3, 0.7, 1.0> }
  hierarchy off sturm
  pigment { rgb <0.3, 0.8, 0.4> } finish { ambient 0.1 diffuse 0.65 specular 0.5 roughness 0.02 reflection 0.15 }
  rotate <0, 25, 8> translate <0.2, 0, 0.5> }
// glass blob with an interior: refraction through Inside / container state
blob { threshold 0.55
  sphere { <3.0, 1.1, -0.5>, 1.4, 1.0 } sphere { <3.9, 1.7, 0.2>, 1.2, 1.0 } sphere { <2.6, 2.1, 0.4>, 1.0, 1.0 }
  cylinder { <2.4, 0.4, -1.2>, <4.2, 0.6, 0.8>, 0.6, 0.8 }
  pigment { rgbf <0.8, 0.9, 1.0, 0.75> } finish { ambient 0.02 diffuse 0.3 specular 0.6 roughness 0.01 reflection 0.1 } interior { ior 1.45 } }
// many small components (deeper bounding-sphere tree), transformed as a whole
blob { threshold 0.4
  #declare I = 0;
  #while (I < 24)
    sphere { <cos(I * 0.7) * (0.6 + 0.05 * I), 0.25 + 0.12 * I, sin(I * 0.7) * (0.6 + 0.05 * I)>, 0.55, 1.0 }
    #declare I = I + 1;
  #end
  pigment { rgb <0.85, 0.75, 0.2> } finish { ambient 0.1 diffuse 0.7 phong 0.3 }
  scale <0.8, 1.0, 0.8> translate <0.5, 0, -3.5> }
