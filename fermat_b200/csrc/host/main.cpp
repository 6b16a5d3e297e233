// main.cpp — headless batch front end with the reference's command line (reference src/main.cu:100-218):
//   fermat_pt -pt -i scene.{fa,obj,fbs} [-r W H] [-c camera.txt] [-bounces N] [-passes P] [-o out] [-device D]
// Renders passes 0..P inclusive (the reference's loop is inclusive: `-passes 1023` = 1024 spp, main.cu:167),
// writes <out>.tga (to_rgba on the device: exposure, c/(1+c), gamma — src/renderer.cu:83-282; `-filtered` runs the EAW
// denoiser first and saves FILTERED_C, src/renderer.cu:1099-1160) and
// <out>.pfm (linear COMPOSITED_C), prints Msamples/s.
//
//   -gpus N [-gather-every K]   one process per GPU (forked here, devices 0..N-1): the frame is tile-sharded over the N processes
//                               (each runs with `-shard r N -device r`), which join one NCCL communicator; every K-th pass (default 1:
//                               every frame) and after the last one the frame is assembled on rank 0 (RenderingContext::
//                               gather_channel_async), which writes the output. The image is the 1-GPU image bit for bit.
#include "rendering_context.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/wait.h>
#include <vector>
#include <chrono>

static int int_option(int argc, char** argv, const char* name, int def)
{
	for (int i = 0; i + 1 < argc; ++i) if (strcmp(argv[i], name) == 0) def = atoi(argv[i + 1]);
	return def;
}

int main(int argc, char** argv)
{
	// ---- -gpus N: fork one process per GPU BEFORE anything touches CUDA; rank 0 (the parent) creates the NCCL id and pipes it to the others
	const int n_gpus = int_option(argc, argv, "-gpus", 1), gather_every = std::max(1, int_option(argc, argv, "-gather-every", 1));
	int rank = 0;
	unsigned char nccl_id[128];
	std::vector<int> id_pipes;          // rank 0: write ends, one per child
	std::vector<pid_t> children;
	int id_read_fd = -1;
	if (n_gpus > 1)
	{
		for (int r = 1; r < n_gpus; ++r)
		{
			int fd[2];
			if (pipe(fd) != 0) { perror("pipe"); return 1; }
			const pid_t pid = fork();
			if (pid < 0) { perror("fork"); return 1; }
			if (pid == 0)
			{
				rank = r; id_read_fd = fd[0]; close(fd[1]);
				for (size_t k = 0; k < id_pipes.size(); ++k) close(id_pipes[k]);
				id_pipes.clear(); children.clear();
				break;
			}
			close(fd[0]); id_pipes.push_back(fd[1]); children.push_back(pid);
		}
	}
	int status = 0;
	try
	{
		std::vector<std::string> extra;
		if (n_gpus > 1) { extra = { "-shard", std::to_string(rank), std::to_string(n_gpus), "-device", std::to_string(rank) }; }
		std::vector<char*> args(argv, argv + argc);
		for (size_t k = 0; k < extra.size(); ++k) args.push_back(const_cast<char*>(extra[k].c_str()));
		if (n_gpus > 1)
		{
			if (rank == 0)
			{
				fb::Communicator::unique_id(nccl_id);
				for (size_t k = 0; k < id_pipes.size(); ++k) { if (write(id_pipes[k], nccl_id, 128) != 128) throw std::runtime_error("cannot hand the NCCL id to a rank"); close(id_pipes[k]); }
			}
			else
			{
				size_t got = 0;
				while (got < 128) { const ssize_t n = read(id_read_fd, nccl_id + got, 128 - got); if (n <= 0) throw std::runtime_error("rank 0 did not send the NCCL id"); got += (size_t)n; }
				close(id_read_fd);
			}
		}
		RenderingContext rc;
		rc.init((int)args.size(), args.data());
		fb200_scene& s = *rc.scene();
		const int n_passes = s.n_passes;
		const std::string out = s.output_name.empty() ? std::string("output") : s.output_name;
		const uint2 res = rc.res();
		if (n_gpus > 1) rc.comm_init(nccl_id, rank, n_gpus);
		float* pinned = NULL;
		if (n_gpus > 1 && rank == 0) pinned = RenderingContext::alloc_pinned((size_t)res.x * res.y * 16);

		rc.clear();
		const auto t0 = std::chrono::steady_clock::now();
		for (int i = 0; i <= n_passes; ++i)
		{
			rc.render((uint32_t)i);
			if (n_gpus > 1 && (i == n_passes || (i + 1) % gather_every == 0)) rc.gather_channel_async(fb::FB_COMPOSITED_C, 0, i == n_passes ? pinned : NULL);
		}
		rc.synchronize();
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		PathTracer* pt = static_cast<PathTracer*>(rc.renderer());
		const fb::PassTotals tot = pt->totals(rc);
		double totals[2] = { (double)tot.shade_events, (double)tot.shadow_events };
		rc.sum_over_ranks(totals, 2);
		if (rank == 0)
			fprintf(stderr, "\n%d passes in %.3f s on %d GPU%s: %.2f Msamples/s (%.0f shade events, %.0f shadow rays), device %.1f ms\n",
				n_passes + 1, sec, n_gpus, n_gpus > 1 ? "s" : "", totals[0] / sec * 1.0e-6, totals[0], totals[1], pt->device_ms());
		if (rank != 0) return 0;

		std::vector<float> img((size_t)res.x * res.y * 4);
		if (n_gpus > 1)
		{
			memcpy(img.data(), pinned, img.size() * sizeof(float));
			// the assembled frame replaces rank 0's shard in its frame buffer, so that to_rgba below shows the whole image
			rc.adopt_gathered_frame(fb::FB_COMPOSITED_C);
			RenderingContext::free_pinned(pinned);
		}
		else rc.download_channel(fb::FB_COMPOSITED_C, img.data());

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
			if (filtered && n_gpus > 1) { fprintf(stderr, "warning: -filtered needs every channel and the G-buffer on one GPU: ignored with -gpus %d\n", n_gpus); filtered = false; }
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
	}
	catch (const std::exception& e)
	{
		// errors are reported and the process exits with a failure code (reference: fprintf + exit(1))
		fprintf(stderr, "error%s: %s\n", n_gpus > 1 ? (" (rank " + std::to_string(rank) + ")").c_str() : "", e.what());
		status = 1;
	}
	for (size_t k = 0; k < children.size(); ++k) { int st = 0; waitpid(children[k], &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) status = 1; }
	return status;
}
