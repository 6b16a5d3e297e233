// main.cpp — headless batch front end with the reference's command line (reference src/main.cu:100-218):
//   fermat_pt -pt -i scene.{fa,obj,fbs} [-r W H] [-c camera.txt] [-bounces N] [-passes P] [-o out] [-device D]
// Renders passes 0..P inclusive (the reference's loop is inclusive: `-passes 1023` = 1024 spp, main.cu:167),
// writes <out>.tga (to_rgba on the device: exposure, c/(1+c), gamma — src/renderer.cu:83-282; `-filtered` runs the EAW
// denoiser first and saves FILTERED_C, src/renderer.cu:1099-1160) and
// <out>.pfm (linear COMPOSITED_C), prints Msamples/s.
#include "rendering_context.h"
#include <stdio.h>
#include <string.h>
#include <vector>
#include <chrono>

int main(int argc, char** argv)
{
	try
	{
		RenderingContext rc;
		rc.init(argc, argv);
		fb200_scene& s = *rc.scene();
		const int n_passes = s.n_passes;
		const std::string out = s.output_name.empty() ? std::string("output") : s.output_name;

		rc.clear();
		const auto t0 = std::chrono::steady_clock::now();
		for (int i = 0; i <= n_passes; ++i) rc.render((uint32_t)i);
		rc.synchronize();
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		PathTracer* pt = static_cast<PathTracer*>(rc.renderer());
		const fb::PassTotals tot = pt->totals(rc);
		fprintf(stderr, "\n%d passes in %.3f s : %.2f Msamples/s (%llu shade events, %llu shadow rays), device %.1f ms\n",
			n_passes + 1, sec, tot.shade_events / sec * 1.0e-6, (unsigned long long)tot.shade_events, (unsigned long long)tot.shadow_events, pt->device_ms());

		const uint2 res = rc.res();
		std::vector<float> img((size_t)res.x * res.y * 4);
		rc.download_channel(fb::FB_COMPOSITED_C, img.data());

		{
			FILE* f = fopen((out + ".pfm").c_str(), "wb");
			if (f)
			{
				fprintf(f, "PF\n%u %u\n-1.0\n", res.x, res.y);
				for (uint32_t y = 0; y < res.y; ++y)        // our row 0 is the bottom row, as PFM expects
					for (uint32_t x = 0; x < res.x; ++x) fwrite(&img[((size_t)y * res.x + x) * 4], 4, 3, f);
				fclose(f);
			}
		}
		{
			// RenderingContext::render's tail (src/renderer.cu:1045-1049): optional EAW filter, then to_rgba; main.cu:171-183 saves it
			bool filtered = false;
			for (int i = 0; i < argc; ++i) if (strcmp(argv[i], "-filtered") == 0) filtered = true;
			if (filtered) rc.filter((uint32_t)n_passes);
			std::vector<uint8_t> rgba((size_t)res.x * res.y * 4);
			rc.to_rgba(filtered ? fb::SHADING_FILTERED : fb::SHADING_SHADED, rgba.data());
			if (fb200_write_tga((out + ".tga").c_str(), res.x, res.y, rgba.data()) != 0) fprintf(stderr, "warning: %s\n", fb200_last_error());
		}
		for (int i = 0; i + 1 < argc; ++i)
			if (strcmp(argv[i], "-benchmark") == 0)
			{
				FILE* f = fopen(argv[i + 1], "w");
				if (f) { rc.renderer()->dump_speed_stats(f); fclose(f); }
			}
		return 0;
	}
	catch (const std::exception& e)
	{
		// errors are reported and the process exits with a failure code (reference: fprintf + exit(1))
		fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
}
