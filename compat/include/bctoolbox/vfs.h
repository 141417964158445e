/* compat shim: the slice of bctoolbox's virtual file system the reference's WAV reader / audiodiff.c use, over POSIX
 * file descriptors. Test infrastructure (oracle/_ref only). */
#ifndef COMPAT_BCTBX_VFS_H
#define COMPAT_BCTBX_VFS_H
#include <fcntl.h>
#include <stdint.h>
#include <sys/types.h>
#ifdef __cplusplus
extern "C" {
#endif
#define BCTBX_VFS_OK 0
#define BCTBX_VFS_ERROR (-255)
typedef struct bctbx_vfs_t bctbx_vfs_t;
typedef struct bctbx_vfs_file_t { /* audiodiff.c:107 moves `offset` by hand: read2 / write2 work at `offset` and advance it */
	int fd;
	off_t offset;
} bctbx_vfs_file_t;
#ifndef BCTBX_EWOULDBLOCK
#define BCTBX_EWOULDBLOCK EWOULDBLOCK
#endif
bctbx_vfs_t *bctbx_vfs_get_default(void);
bctbx_vfs_file_t *bctbx_file_open(bctbx_vfs_t *vfs, const char *path, const char *mode);
bctbx_vfs_file_t *bctbx_file_open2(bctbx_vfs_t *vfs, const char *path, int openflags);
int64_t bctbx_file_size(bctbx_vfs_file_t *f);
int bctbx_file_close(bctbx_vfs_file_t *f);
ssize_t bctbx_file_read(bctbx_vfs_file_t *f, void *buf, size_t count, off_t offset);
ssize_t bctbx_file_read2(bctbx_vfs_file_t *f, void *buf, size_t count);
ssize_t bctbx_file_write(bctbx_vfs_file_t *f, const void *buf, size_t count, off_t offset);
ssize_t bctbx_file_write2(bctbx_vfs_file_t *f, const void *buf, size_t count);
off_t bctbx_file_seek(bctbx_vfs_file_t *f, off_t offset, int whence);
int bctbx_file_truncate(bctbx_vfs_file_t *f, int64_t size);
int bctbx_file_sync(bctbx_vfs_file_t *f);
#ifdef __cplusplus
}
#endif
#endif
