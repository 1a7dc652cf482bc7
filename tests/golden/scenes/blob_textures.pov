// blobs whose components carry their own textures: Blob::Determine_Textures (blob.cpp:2768-2880) blends them by field contribution;
// one component texture is a texture_map, one blob is transparent (weighted filter colours on the shadow rays)
#version 3.7;
global_settings { assumed_gamma 1 max_trace_level 5 }
background { rgb <0.15, 0.18, 0.3> }
camera { location <0.02, 2.6, -7> look_at <0, 0.9, 0> angle 40 right x*16/9 }
light_source { <5, 9, -6> rgb 1 }
light_source { <-5, 4, -4> rgb <0.3, 0.3, 0.4> }
plane { y, -0.0078125 pigment { checker rgb 0.85, rgb 0.35 } finish { ambient 0.1 diffuse 0.7 } }
blob {
  threshold 0.6
  sphere { <-0.9, 0, 0>, 1.1, 1 texture { pigment { rgb <0.9, 0.2, 0.15> } finish { ambient 0.1 diffuse 0.6 phong 0.6 } } }
  sphere { <0.1, 0.5, 0>, 1.0, 1 texture { pigment { rgb <0.2, 0.8, 0.25> } finish { ambient 0.1 diffuse 0.6 } } }
  sphere { <0.9, -0.1, 0.2>, 1.1, 1 texture { gradient y texture_map { [0.2 pigment { rgb <0.15, 0.25, 0.9> } finish { diffuse 0.6 } ] [0.8 pigment { rgb <0.95, 0.9, 0.2> } finish { diffuse 0.6 specular 0.5 } ] } scale 0.8 } }
  cylinder { <-0.9, 0, 0>, <0.9, 0, 0.2>, 0.45, 0.8 }
  pigment { rgb 0.8 } finish { ambient 0.1 diffuse 0.6 }
  translate <-1.4, 1.0, 0.4>
}
blob {
  threshold 0.5
  sphere { <0, 0, 0>, 1.0, 1 texture { pigment { rgbf <1.0, 0.4, 0.3, 0.75> } finish { ambient 0.05 diffuse 0.3 specular 0.5 } } }
  sphere { <0.8, 0.4, 0>, 0.9, 1 texture { pigment { rgbf <0.3, 0.5, 1.0, 0.75> } finish { ambient 0.05 diffuse 0.3 specular 0.5 } } }
  sphere { <0.4, -0.3, 0.5>, 0.7, -0.4 }
  pigment { rgbf <0.9, 0.9, 0.9, 0.6> } finish { diffuse 0.3 }
  interior { ior 1.3 }
  translate <1.6, 0.9, -0.6>
}
