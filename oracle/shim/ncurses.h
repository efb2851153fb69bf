/*
 * oracle/shim/ncurses.h -- TEST INFRASTRUCTURE, not product code.
 * No-op stand-in for the ncurses calls in the Barcode reference
 * (curses_funcs.cc:13-98, wprintw/wrefresh in HMC.cc:118,135,269,285): the
 * oracle harness has no terminal UI.
 */
#ifndef BARCODE_ORACLE_SHIM_NCURSES_H
#define BARCODE_ORACLE_SHIM_NCURSES_H
typedef struct shim_window { int unused; } WINDOW;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define COLOR_BLACK 0
#define COLOR_BLUE 4
#define COLOR_YELLOW 3
#define COLOR_WHITE 7
#define COLOR_CYAN 6
#define A_BOLD 0
#define COLOR_PAIR(n) (n)
#ifdef __cplusplus
extern "C" {
#endif
extern WINDOW *stdscr;
WINDOW *initscr(void);
int endwin(void);
int isendwin(void);
int refresh(void);
int start_color(void);
int cbreak(void);
int noecho(void);
int intrflush(WINDOW *, int);
int keypad(WINDOW *, int);
int init_pair(short, short, short);
WINDOW *newwin(int, int, int, int);
int wattron(WINDOW *, int);
int wbkgd(WINDOW *, unsigned);
int scrollok(WINDOW *, int);
int wprintw(WINDOW *, const char *, ...);
int wrefresh(WINDOW *);
#ifdef __cplusplus
}
#endif
#define getmaxyx(win, y, x) ((y) = 24, (x) = 80)
#endif
